"""ctypes binding of include/edsparser_b200.h.

Plumbing for tests/, bench.py and __graft_entry__ only: the product's host side is the C++17 layer
under edsparser_b200/host/ (same signatures as the reference's transforms/ headers). The library
has no CPU path: `load()` raises if libedsparser_b200.so has not been built, and every transform
raises EdsError(status=EDS_ERR_CUDA) when no CUDA device is usable.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_SO = os.environ.get("EDSB_LIBRARY") or os.path.join(HERE, "libedsparser_b200.so")  # (another build of the same sources)

EDS_OK = 0
EDS_ERR_INVALID_ARGUMENT = 1
EDS_ERR_RUNTIME = 2
EDS_ERR_OUT_OF_RANGE = 3
EDS_ERR_CUDA = 4
EDS_ERR_BAD_MSA = 5
EDS_ERR_BUDGET = 6
EDS_ERR_HALO = 7
EDS_ERR_BAD_VCF = 8

EXPORTS = [
    "eds_last_error", "eds_version", "eds_ctx_create", "eds_ctx_destroy", "eds_ctx_synchronize",
    "eds_ctx_set_tuning", "eds_ctx_set_profiling", "eds_ctx_kernel_times", "eds_msa_index_host",
    "eds_msa_index_free", "eds_msa_transform_device", "eds_msa_transform_host", "eds_msa_transform_host_view",
    "eds_msa_conserved_bits",
    "eds_msa_synth_device", "eds_msa_synth_free", "eds_buffer_to_host", "eds_buffer_to_host_view", "eds_buffer_free_host", "eds_leds_merge_host", "eds_leds_merge_host_view", "eds_is_leds_host",
    "eds_vcf_transform_host", "eds_vcf_transform_host_view", "eds_vcf_transform_device", "eds_device_upload", "eds_device_free",
    "eds_group_create", "eds_group_destroy", "eds_group_size", "eds_group_ctx", "eds_group_msa_transform_host",
    "eds_group_msa_transform_fd", "eds_nccl_unique_id", "eds_comm_create", "eds_comm_destroy", "eds_comm_post",
    "eds_comm_offsets", "eds_comm_flush", "eds_parse_host", "eds_parsed_free", "eds_merge_adjacent_host",
    "eds_group_leds_merge_host", "eds_group_vcf_transform_host", "eds_group_vcf_transform_host_view", "eds_leds_merge_device_in", "eds_genrandomeds_device",
]


class EdsError(Exception):
    def __init__(self, status, message):
        super().__init__(f"[eds_status {status}] {message}")
        self.status = status
        self.message = message


class MsaIndex(ctypes.Structure):
    _fields_ = [("row_start", ctypes.POINTER(ctypes.c_uint64)), ("n_rows", ctypes.c_uint32),
                ("n_cols", ctypes.c_uint64), ("line_width", ctypes.c_uint32), ("row_bytes", ctypes.c_uint64)]


class MsaView(ctypes.Structure):
    _fields_ = [("text", ctypes.c_void_p), ("text_bytes", ctypes.c_uint64),
                ("row_start", ctypes.POINTER(ctypes.c_uint64)), ("n_rows", ctypes.c_uint32),
                ("line_width", ctypes.c_uint32), ("total_cols", ctypes.c_uint64), ("col_begin", ctypes.c_uint64),
                ("col_count", ctypes.c_uint64), ("own_begin", ctypes.c_uint64), ("own_end", ctypes.c_uint64)]


class Buffer(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("bytes", ctypes.c_uint64)]


class MsaStats(ctypes.Structure):
    _fields_ = [("n_variable_cols", ctypes.c_uint64), ("n_runs", ctypes.c_uint64), ("n_symbols", ctypes.c_uint64),
                ("n_variable", ctypes.c_uint64), ("n_alternatives", ctypes.c_uint64),
                ("first_open_col", ctypes.c_uint64), ("eds_bytes", ctypes.c_uint64), ("seds_bytes", ctypes.c_uint64),
                ("eds_lead_bytes", ctypes.c_uint64), ("tail_open_common", ctypes.c_uint32),
                ("gpu_launches", ctypes.c_uint32), ("retries", ctypes.c_uint32), ("n_hashed_symbols", ctypes.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class VcfStats(ctypes.Structure):
    _fields_ = [(k, ctypes.c_uint64) for k in (
        "total_variants", "processed_variants", "skipped_malformed", "skipped_unsupported_sv", "variant_groups",
        "n_lines", "n_bases", "n_alleles", "n_haplotype_slots", "n_samples_max", "eds_bytes", "seds_bytes")] + [
        (k, ctypes.c_uint32) for k in ("gpu_launches", "host_sorted", "retries", "leds_rounds")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Library:
    """One loaded copy of the C ABI."""

    def __init__(self, path=PRODUCT_SO):
        if not os.path.exists(path):
            raise ImportError(f"{path} is not built (run `make` or __graft_entry__.build()); there is no CPU fallback")
        self.path = path
        L = ctypes.CDLL(path)
        vp, u8p = ctypes.c_void_p, ctypes.c_char_p
        u32, u64, i32 = ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
        P = ctypes.POINTER
        L.eds_last_error.restype = ctypes.c_char_p
        L.eds_version.restype = ctypes.c_char_p
        L.eds_ctx_create.argtypes = [i32, vp, P(vp)]
        L.eds_ctx_destroy.argtypes = [vp]
        L.eds_ctx_destroy.restype = None
        L.eds_ctx_synchronize.argtypes = [vp]
        L.eds_ctx_set_tuning.argtypes = [vp, u32, u32]
        L.eds_ctx_set_profiling.argtypes = [vp, i32]
        L.eds_ctx_kernel_times.argtypes = [vp, P(ctypes.c_char_p), P(ctypes.c_float), u32]
        L.eds_ctx_kernel_times.restype = u32
        L.eds_msa_index_host.argtypes = [vp, u64, P(MsaIndex)]
        L.eds_msa_index_free.argtypes = [P(MsaIndex)]
        L.eds_msa_index_free.restype = None
        L.eds_msa_transform_device.argtypes = [vp, P(MsaView), u32, i32, P(Buffer), P(Buffer), P(MsaStats)]
        L.eds_msa_transform_host.argtypes = [vp, vp, u64, u32, i32, P(Buffer), P(Buffer), P(MsaStats)]
        L.eds_msa_transform_host_view.argtypes = L.eds_msa_transform_host.argtypes
        L.eds_msa_conserved_bits.argtypes = [vp, P(MsaView), vp, u64]
        L.eds_msa_synth_device.argtypes = [vp, u32, u64, u32, u64, u64, u64, u32, P(MsaView)]
        L.eds_msa_synth_free.argtypes = [vp]
        L.eds_msa_synth_free.restype = None
        L.eds_buffer_to_host.argtypes = [vp, P(Buffer), P(Buffer)]
        L.eds_buffer_to_host_view.argtypes = [vp, i32, P(Buffer), P(Buffer)]
        L.eds_buffer_free_host.argtypes = [P(Buffer)]
        L.eds_buffer_free_host.restype = None
        L.eds_leds_merge_host.argtypes = [vp, vp, u64, vp, u64, u32, i32, u64, P(Buffer), P(Buffer), P(u32)]
        L.eds_leds_merge_host_view.argtypes = L.eds_leds_merge_host.argtypes
        L.eds_is_leds_host.argtypes = [vp, vp, u64, u32, P(i32)]
        L.eds_vcf_transform_host.argtypes = [vp, vp, u64, vp, u64, u32, P(Buffer), P(Buffer), P(VcfStats),
                                             P(P(ctypes.c_uint64)), P(u64)]
        L.eds_vcf_transform_host_view.argtypes = L.eds_vcf_transform_host.argtypes
        L.eds_vcf_transform_device.argtypes = [vp, vp, u64, vp, u64, P(Buffer), P(Buffer), P(VcfStats)]
        L.eds_device_upload.argtypes = [vp, vp, u64, P(vp)]
        L.eds_device_free.argtypes = [vp, vp]
        L.eds_device_free.restype = None
        L.eds_group_create.argtypes = [P(i32), i32, P(vp)]
        L.eds_group_destroy.argtypes = [vp]
        L.eds_group_destroy.restype = None
        L.eds_group_size.argtypes = [vp]
        L.eds_group_ctx.argtypes = [vp, i32]
        L.eds_group_ctx.restype = vp
        L.eds_group_msa_transform_host.argtypes = [vp, vp, u64, u32, i32, u64, P(Buffer), P(Buffer), P(MsaStats)]
        L.eds_group_msa_transform_fd.argtypes = [vp, vp, u64, u32, i32, u64, i32, i32, P(u64), P(MsaStats)]
        L.eds_group_leds_merge_host.argtypes = [vp, vp, u64, vp, u64, u32, i32, P(Buffer), P(Buffer), P(u32), P(u32)]
        L.eds_group_vcf_transform_host.argtypes = L.eds_vcf_transform_host.argtypes + [P(u32)]
        L.eds_group_vcf_transform_host_view.argtypes = L.eds_group_vcf_transform_host.argtypes
        L.eds_leds_merge_device_in.argtypes = [vp, vp, u64, vp, u64, u32, i32, u64, P(Buffer), P(Buffer), P(u32)]
        L.eds_genrandomeds_device.argtypes = [vp, u64, u32, u32, u64, P(Buffer), P(Buffer)]
        L.eds_nccl_unique_id.argtypes = [vp]
        L.eds_comm_create.argtypes = [vp, vp, i32, i32, P(vp)]
        L.eds_comm_destroy.argtypes = [vp]
        L.eds_comm_destroy.restype = None
        L.eds_comm_post.argtypes = [vp, u64, u64]
        L.eds_comm_offsets.argtypes = [vp, P(u64)]
        L.eds_comm_flush.argtypes = [vp]
        self.L = L
        _ = u8p

    def version(self):
        return self.L.eds_version().decode()

    def check(self, rc):
        if rc != EDS_OK:
            raise EdsError(rc, self.L.eds_last_error().decode("latin-1"))

    def context(self, device=0, stream=None):
        return Context(self, device, stream)

    def group(self, devices):
        return Group(self, devices)

    def nccl_unique_id(self):
        buf = ctypes.create_string_buffer(128)
        self.check(self.L.eds_nccl_unique_id(buf))
        return buf.raw


class Group:
    """eds_group: one process, N devices, column-sharded msa2eds with the NCCL exchange inside the library."""

    def __init__(self, lib, devices):
        self.lib = lib
        devices = list(devices)
        arr = (ctypes.c_int * len(devices))(*devices)
        h = ctypes.c_void_p()
        lib.check(lib.L.eds_group_create(arr, len(devices), ctypes.byref(h)))
        self.handle, self.n = h, len(devices)

    def close(self):
        if self.handle:
            self.lib.L.eds_group_destroy(self.handle)
            self.handle = None

    def msa_transform_host(self, file_bytes, l, leds=None, halo=0):
        leds = (1 if l > 0 else 0) if leds is None else leds
        e, s = Buffer(), Buffer()
        st = (MsaStats * self.n)()
        src = ctypes.c_char_p(bytes(file_bytes))
        self.lib.check(self.lib.L.eds_group_msa_transform_host(self.handle, src, len(file_bytes), l, leds, halo, ctypes.byref(e),
                                                               ctypes.byref(s), st))
        return _host_bytes(self.lib, e), _host_bytes(self.lib, s), [x.as_dict() for x in st]  # (_host_bytes frees)

    def leds_merge_host(self, eds, seds, l, compact=True):
        """eds2leds over the group's devices: (l-EDS bytes, SEDS bytes or b"", rounds, shards used)."""
        e, s = Buffer(), Buffer()
        rounds, used = ctypes.c_uint32(), ctypes.c_uint32()
        se = ctypes.c_char_p(bytes(eds))
        ss = ctypes.c_char_p(bytes(seds)) if seds is not None else None
        self.lib.check(self.lib.L.eds_group_leds_merge_host(self.handle, se, len(eds), ss, len(seds) if seds is not None else 0, l,
                                                            1 if compact else 0, ctypes.byref(e), ctypes.byref(s), ctypes.byref(rounds),
                                                            ctypes.byref(used)))
        return _host_bytes(self.lib, e), _host_bytes(self.lib, s), rounds.value, used.value

    def vcf_transform_host(self, vcf, fasta, l=0):
        """vcf2eds over the group's devices: (eds, seds, stats dict, SV line offsets, shards used)."""
        va, vn, k1 = _as_pointer(vcf)
        fa, fn, k2 = _as_pointer(fasta)
        e, s, st = Buffer(), Buffer(), VcfStats()
        sv, nsv, used = ctypes.POINTER(ctypes.c_uint64)(), ctypes.c_uint64(), ctypes.c_uint32()
        self.lib.check(self.lib.L.eds_group_vcf_transform_host(self.handle, va, vn, fa, fn, l, ctypes.byref(e), ctypes.byref(s),
                                                               ctypes.byref(st), ctypes.byref(sv), ctypes.byref(nsv), ctypes.byref(used)))
        del k1, k2
        lines = [int(sv[i]) for i in range(nsv.value)]
        if nsv.value:
            ctypes.CDLL(None).free(sv)
        return _host_bytes(self.lib, e), _host_bytes(self.lib, s), st.as_dict(), lines, used.value

    def vcf_transform_host_view(self, vcf, fasta, l=0):
        """Like vcf_transform_host through eds_group_vcf_transform_host_view (pinned views, copied into bytes here)."""
        va, vn, k1 = _as_pointer(vcf)
        fa, fn, k2 = _as_pointer(fasta)
        e, s, st, used = Buffer(), Buffer(), VcfStats(), ctypes.c_uint32()
        self.lib.check(self.lib.L.eds_group_vcf_transform_host_view(self.handle, va, vn, fa, fn, l, ctypes.byref(e), ctypes.byref(s),
                                                                    ctypes.byref(st), None, None, ctypes.byref(used)))
        del k1, k2
        eds = bytes((ctypes.c_ubyte * e.bytes).from_address(e.data)) if e.bytes else b""
        seds = bytes((ctypes.c_ubyte * s.bytes).from_address(s.data)) if s.bytes else b""
        return eds, seds, st.as_dict(), used.value

    def vcf_transform_host_view_raw(self, vcf_addr, vcf_n, fa_addr, fa_n, l=0):
        """The bare C call of the view form: ((eds bytes, seds bytes), stats, shards used); nothing to free."""
        e, s, st, used = Buffer(), Buffer(), VcfStats(), ctypes.c_uint32()
        self.lib.check(self.lib.L.eds_group_vcf_transform_host_view(self.handle, vcf_addr, vcf_n, fa_addr, fa_n, l, ctypes.byref(e),
                                                                    ctypes.byref(s), ctypes.byref(st), None, None, ctypes.byref(used)))
        return (int(e.bytes), int(s.bytes)), st.as_dict(), used.value

    def vcf_transform_host_raw(self, vcf_addr, vcf_n, fa_addr, fa_n, l=0, keep=False):
        """The bare C call on (address, length) pairs: ((eds bytes, seds bytes), stats, shards used[, eds, seds])."""
        e, s, st, used = Buffer(), Buffer(), VcfStats(), ctypes.c_uint32()
        self.lib.check(self.lib.L.eds_group_vcf_transform_host(self.handle, vcf_addr, vcf_n, fa_addr, fa_n, l, ctypes.byref(e),
                                                               ctypes.byref(s), ctypes.byref(st), None, None, ctypes.byref(used)))
        sizes = (int(e.bytes), int(s.bytes))
        if keep:
            return sizes, st.as_dict(), used.value, _host_bytes(self.lib, e), _host_bytes(self.lib, s)
        self.lib.L.eds_buffer_free_host(ctypes.byref(e))
        self.lib.L.eds_buffer_free_host(ctypes.byref(s))
        return sizes, st.as_dict(), used.value

    def msa_transform_files(self, file_bytes, l, eds_path, seds_path, leds=None, halo=0):
        leds = (1 if l > 0 else 0) if leds is None else leds
        fe = os.open(eds_path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
        fs = os.open(seds_path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
        try:
            tot = (ctypes.c_uint64 * 2)()
            src = ctypes.c_char_p(bytes(file_bytes))
            self.lib.check(self.lib.L.eds_group_msa_transform_fd(self.handle, src, len(file_bytes), l, leds, halo, fe, fs, tot, None))
            return int(tot[0]), int(tot[1])
        finally:
            os.close(fe)
            os.close(fs)


class Comm:
    """eds_comm: one process per GPU; the byte-count all-gather behind each transform, NCCL called from the library."""

    def __init__(self, ctx, unique_id, rank, world):
        self.lib = ctx.lib
        h = ctypes.c_void_p()
        idbuf = ctypes.create_string_buffer(bytes(unique_id), 128) if unique_id is not None else None
        self.lib.check(self.lib.L.eds_comm_create(ctx.handle, idbuf, rank, world, ctypes.byref(h)))
        self.handle = h
        self._out = (ctypes.c_uint64 * 4)()

    def post(self, eds_bytes, seds_bytes):
        rc = self.lib.L.eds_comm_post(self.handle, eds_bytes, seds_bytes)
        if rc:
            self.lib.check(rc)

    def flush(self):
        self.lib.check(self.lib.L.eds_comm_flush(self.handle))

    def offsets(self):
        self.lib.check(self.lib.L.eds_comm_offsets(self.handle, self._out))
        return tuple(int(x) for x in self._out)

    def close(self):
        if self.handle:
            self.lib.L.eds_comm_destroy(self.handle)
            self.handle = None


def _host_bytes(lib, buf):
    # (ctypes.string_at takes a C int: outputs past 2 GiB go through an array view)
    data = bytes((ctypes.c_ubyte * buf.bytes).from_address(buf.data)) if buf.data and buf.bytes else b""
    lib.L.eds_buffer_free_host(ctypes.byref(buf))
    return data


def _as_pointer(data):
    """(address, length, keepalive) of bytes / bytearray / numpy / torch host data."""
    if isinstance(data, bytes):  # read-only input: the object's own buffer, no copy
        ref = ctypes.c_char_p(data)
        return ctypes.cast(ref, ctypes.c_void_p).value or 0, len(data), (ref, data)
    if isinstance(data, bytearray):
        arr = (ctypes.c_char * len(data)).from_buffer(data)
        return ctypes.addressof(arr), len(data), arr
    if hasattr(data, "data_ptr"):  # torch tensor (host, uint8)
        return data.data_ptr(), data.numel() * data.element_size(), data
    if hasattr(data, "ctypes"):  # numpy
        return data.ctypes.data, data.nbytes, data
    raise TypeError(type(data))


class Context:
    def __init__(self, lib, device=0, stream=None):
        self.lib = lib
        self.handle = ctypes.c_void_p()
        # stream None: the library makes its own stream. An integer is a cudaStream_t; torch reports the
        # legacy default stream as 0, which the ABI spells cudaStreamLegacy (0x1) because NULL means "own".
        arg = None if stream is None else ctypes.c_void_p(stream if stream else 1)
        lib.check(lib.L.eds_ctx_create(device, arg, ctypes.byref(self.handle)))

    def close(self):
        if self.handle:
            self.lib.L.eds_ctx_destroy(self.handle)
            self.handle = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_tuning(self, partitions=0, scan_blocks_per_sm=0):
        self.lib.check(self.lib.L.eds_ctx_set_tuning(self.handle, partitions, scan_blocks_per_sm))

    def set_profiling(self, on):
        self.lib.check(self.lib.L.eds_ctx_set_profiling(self.handle, 1 if on else 0))

    def kernel_times(self):
        cap = 512
        names = (ctypes.c_char_p * cap)()
        ms = (ctypes.c_float * cap)()
        n = min(cap, self.lib.L.eds_ctx_kernel_times(self.handle, names, ms, cap))
        return [(names[i].decode(), float(ms[i])) for i in range(n) if names[i]]

    def synchronize(self):
        self.lib.check(self.lib.L.eds_ctx_synchronize(self.handle))

    # -- MSA ------------------------------------------------------------------------------
    def msa_index(self, text):
        addr, n, keep = _as_pointer(text)
        idx = MsaIndex()
        self.lib.check(self.lib.L.eds_msa_index_host(addr, n, ctypes.byref(idx)))
        out = {"row_start": [idx.row_start[i] for i in range(idx.n_rows)], "n_rows": idx.n_rows,
               "n_cols": idx.n_cols, "line_width": idx.line_width, "row_bytes": idx.row_bytes}
        self.lib.L.eds_msa_index_free(ctypes.byref(idx))
        del keep
        return out

    def genrandomeds_device(self, ref_size, variability_ppm=100_000, paths=4, seed=1):
        """genrandomeds-shaped EDS + SEDS text generated in device memory: (eds Buffer, seds Buffer), owned by the ctx."""
        e, s = Buffer(), Buffer()
        self.lib.check(self.lib.L.eds_genrandomeds_device(self.handle, ref_size, variability_ppm, paths, seed, ctypes.byref(e),
                                                          ctypes.byref(s)))
        return e, s

    def leds_merge_device_in(self, eds_buf, seds_buf, l, compact=True, max_output_bytes=0):
        """eds2leds on text already in device memory: (l-EDS bytes, SEDS bytes, rounds)."""
        o, so, rounds = Buffer(), Buffer(), ctypes.c_uint32()
        self.lib.check(self.lib.L.eds_leds_merge_device_in(self.handle, eds_buf.data, eds_buf.bytes, seds_buf.data if seds_buf else None,
                                                           seds_buf.bytes if seds_buf else 0, l, 1 if compact else 0, max_output_bytes,
                                                           ctypes.byref(o), ctypes.byref(so), ctypes.byref(rounds)))
        return _host_bytes(self.lib, o), _host_bytes(self.lib, so), rounds.value

    def msa_transform_host(self, text, l=0, leds=None):
        """bytes of a .msa file -> (eds bytes, seds bytes, stats dict); l == 0 and leds None -> plain EDS."""
        if leds is None:
            leds = l > 0
        addr, n, keep = _as_pointer(text)
        e, s, st = Buffer(), Buffer(), MsaStats()
        self.lib.check(self.lib.L.eds_msa_transform_host(self.handle, addr, n, l, 1 if leds else 0, ctypes.byref(e),
                                                         ctypes.byref(s), ctypes.byref(st)))
        del keep
        return _host_bytes(self.lib, e), _host_bytes(self.lib, s), st.as_dict()

    def msa_transform_host_view_raw(self, addr, n, l, leds=True):
        """The bare C call of the view form (results stay in pinned memory kept by the context): sizes and stats."""
        e, s, st = Buffer(), Buffer(), MsaStats()
        self.lib.check(self.lib.L.eds_msa_transform_host_view(self.handle, addr, n, l, 1 if leds else 0, ctypes.byref(e),
                                                              ctypes.byref(s), ctypes.byref(st)))
        return (int(e.bytes), int(s.bytes)), st.as_dict()

    def msa_transform_host_view(self, text, l=0, leds=None):
        """Like msa_transform_host through eds_msa_transform_host_view (views copied into bytes here)."""
        if leds is None:
            leds = l > 0
        addr, n, keep = _as_pointer(text)
        e, s, st = Buffer(), Buffer(), MsaStats()
        self.lib.check(self.lib.L.eds_msa_transform_host_view(self.handle, addr, n, l, 1 if leds else 0, ctypes.byref(e),
                                                              ctypes.byref(s), ctypes.byref(st)))
        del keep
        eds = bytes((ctypes.c_ubyte * e.bytes).from_address(e.data)) if e.bytes else b""
        seds = bytes((ctypes.c_ubyte * s.bytes).from_address(s.data)) if s.bytes else b""
        return eds, seds, st.as_dict()

    def msa_transform_host_raw(self, addr, n, l, leds=True):
        """The bare C call on (address, length) of host bytes; outputs are freed, only sizes and stats are
        returned (bench.py's end-to-end timing: no Python-side copies inside the timed region)."""
        e, s, st = Buffer(), Buffer(), MsaStats()
        self.lib.check(self.lib.L.eds_msa_transform_host(self.handle, addr, n, l, 1 if leds else 0, ctypes.byref(e),
                                                         ctypes.byref(s), ctypes.byref(st)))
        sizes = (int(e.bytes), int(s.bytes))
        self.lib.L.eds_buffer_free_host(ctypes.byref(e))
        self.lib.L.eds_buffer_free_host(ctypes.byref(s))
        return sizes, st.as_dict()

    def msa_transform_device(self, view, l=0, leds=None):
        """view: MsaView over device memory -> (eds Buffer, seds Buffer, stats dict); buffers stay on the device."""
        if leds is None:
            leds = l > 0
        e, s, st = Buffer(), Buffer(), MsaStats()
        self.lib.check(self.lib.L.eds_msa_transform_device(self.handle, ctypes.byref(view), l, 1 if leds else 0,
                                                           ctypes.byref(e), ctypes.byref(s), ctypes.byref(st)))
        return e, s, st.as_dict()

    def download(self, device_buf):
        """bytes of a device-resident Buffer (an output of msa_transform_device, or any (ptr, n) pair)."""
        if not device_buf.bytes:
            return b""
        h = Buffer()
        self.lib.check(self.lib.L.eds_buffer_to_host(self.handle, ctypes.byref(device_buf), ctypes.byref(h)))
        return _host_bytes(self.lib, h)

    def download_view(self, slot, device_buf):
        """Copy a device-resident Buffer into the context's pinned slot (0 or 1); returns the host Buffer (a view)."""
        h = Buffer()
        self.lib.check(self.lib.L.eds_buffer_to_host_view(self.handle, slot, ctypes.byref(device_buf), ctypes.byref(h)))
        return h

    def msa_conserved_bits(self, view):
        n = (view.col_count + 7) // 8
        out = ctypes.create_string_buffer(n)
        self.lib.check(self.lib.L.eds_msa_conserved_bits(self.handle, ctypes.byref(view), out, n))
        return out.raw

    def msa_synth(self, n_rows, total_cols, line_width=80, col_begin=0, col_count=None, seed=1, variable_ppm=10000):
        if col_count is None:
            col_count = total_cols - col_begin
        v = MsaView()
        self.lib.check(self.lib.L.eds_msa_synth_device(self.handle, n_rows, total_cols, line_width, col_begin,
                                                       col_count, seed, variable_ppm, ctypes.byref(v)))
        return v

    def msa_synth_free(self):
        self.lib.L.eds_msa_synth_free(self.handle)

    # -- l-EDS merge ------------------------------------------------------------------------
    def leds_merge_host(self, eds, seds, l, compact=True, max_output_bytes=0):
        """seds None -> CARTESIAN. Returns (leds bytes, seds bytes or None, rounds)."""
        ea, en, k1 = _as_pointer(eds)
        if seds is not None:
            sa, sn, k2 = _as_pointer(seds)
        else:
            sa, sn, k2 = None, 0, None
        o, so, rounds = Buffer(), Buffer(), ctypes.c_uint32()
        self.lib.check(self.lib.L.eds_leds_merge_host(self.handle, ea, en, sa, sn, l, 1 if compact else 0,
                                                      max_output_bytes, ctypes.byref(o), ctypes.byref(so),
                                                      ctypes.byref(rounds)))
        del k1, k2
        out, sout = _host_bytes(self.lib, o), _host_bytes(self.lib, so)
        return out, (sout if seds is not None else None), rounds.value


    def leds_merge_host_view_raw(self, eds_addr, eds_n, seds_addr, seds_n, l, compact=True, max_output_bytes=0):
        """The bare C call of the view form on (address, length) pairs: ((leds bytes, seds bytes), rounds)."""
        o, so, rounds = Buffer(), Buffer(), ctypes.c_uint32()
        self.lib.check(self.lib.L.eds_leds_merge_host_view(self.handle, eds_addr, eds_n, seds_addr, seds_n, l,
                                                           1 if compact else 0, max_output_bytes, ctypes.byref(o),
                                                           ctypes.byref(so), ctypes.byref(rounds)))
        return (int(o.bytes), int(so.bytes)), rounds.value

    def leds_merge_host_view(self, eds, seds, l, compact=True, max_output_bytes=0):
        """Like leds_merge_host through eds_leds_merge_host_view (views into pinned memory, copied into bytes here)."""
        ea, en, k1 = _as_pointer(eds)
        sa, sn, k2 = _as_pointer(seds) if seds is not None else (None, 0, None)
        o, so, rounds = Buffer(), Buffer(), ctypes.c_uint32()
        self.lib.check(self.lib.L.eds_leds_merge_host_view(self.handle, ea, en, sa, sn, l, 1 if compact else 0,
                                                           max_output_bytes, ctypes.byref(o), ctypes.byref(so),
                                                           ctypes.byref(rounds)))
        del k1, k2
        out = bytes((ctypes.c_ubyte * o.bytes).from_address(o.data)) if o.bytes else b""
        sout = bytes((ctypes.c_ubyte * so.bytes).from_address(so.data)) if so.bytes else b""
        return out, (sout if seds is not None else None), rounds.value

    def is_leds(self, eds, l):
        ea, en, keep = _as_pointer(eds)
        out = ctypes.c_int()
        self.lib.check(self.lib.L.eds_is_leds_host(self.handle, ea, en, l, ctypes.byref(out)))
        del keep
        return bool(out.value)

    # -- VCF front end --------------------------------------------------------------------------
    def vcf_transform_host(self, vcf, fasta, l=0):
        """bytes of a .vcf and a .fa -> (eds, seds, stats dict, offsets of the lines skipped as unsupported SVs)."""
        va, vn, k1 = _as_pointer(vcf)
        fa, fn, k2 = _as_pointer(fasta)
        e, s, st = Buffer(), Buffer(), VcfStats()
        sv, nsv = ctypes.POINTER(ctypes.c_uint64)(), ctypes.c_uint64()
        self.lib.check(self.lib.L.eds_vcf_transform_host(self.handle, va, vn, fa, fn, l, ctypes.byref(e), ctypes.byref(s),
                                                         ctypes.byref(st), ctypes.byref(sv), ctypes.byref(nsv)))
        del k1, k2
        lines = [int(sv[i]) for i in range(nsv.value)]
        if nsv.value:
            ctypes.CDLL(None).free(sv)
        return _host_bytes(self.lib, e), _host_bytes(self.lib, s), st.as_dict(), lines

    def vcf_transform_host_view(self, vcf, fasta, l=0):
        """Like vcf_transform_host, through eds_vcf_transform_host_view (results are views into pinned memory owned by
        the context; copied into bytes here)."""
        va, vn, k1 = _as_pointer(vcf)
        fa, fn, k2 = _as_pointer(fasta)
        e, s, st = Buffer(), Buffer(), VcfStats()
        self.lib.check(self.lib.L.eds_vcf_transform_host_view(self.handle, va, vn, fa, fn, l, ctypes.byref(e), ctypes.byref(s),
                                                              ctypes.byref(st), None, None))
        del k1, k2
        eds = bytes((ctypes.c_ubyte * e.bytes).from_address(e.data)) if e.bytes else b""
        seds = bytes((ctypes.c_ubyte * s.bytes).from_address(s.data)) if s.bytes else b""
        return eds, seds, st.as_dict()

    def vcf_transform_host_view_raw(self, vcf_addr, vcf_n, fa_addr, fa_n, l=0):
        """The bare C call of the view form: (eds bytes, seds bytes), stats; nothing to free."""
        e, s, st = Buffer(), Buffer(), VcfStats()
        self.lib.check(self.lib.L.eds_vcf_transform_host_view(self.handle, vcf_addr, vcf_n, fa_addr, fa_n, l, ctypes.byref(e),
                                                              ctypes.byref(s), ctypes.byref(st), None, None))
        return (int(e.bytes), int(s.bytes)), st.as_dict()

    def vcf_transform_host_raw(self, vcf_addr, vcf_n, fa_addr, fa_n, l=0):
        """The bare C call on (address, length) pairs of host memory; outputs are freed, only sizes and stats are
        returned (tools/bench_vcf.py's host-to-host timing: no Python-side copies inside the timed region)."""
        e, s, st = Buffer(), Buffer(), VcfStats()
        self.lib.check(self.lib.L.eds_vcf_transform_host(self.handle, vcf_addr, vcf_n, fa_addr, fa_n, l, ctypes.byref(e),
                                                         ctypes.byref(s), ctypes.byref(st), None, None))
        sizes = (int(e.bytes), int(s.bytes))
        self.lib.L.eds_buffer_free_host(ctypes.byref(e))
        self.lib.L.eds_buffer_free_host(ctypes.byref(s))
        return sizes, st.as_dict()

    def upload(self, data):
        """host bytes -> Buffer over fresh device memory (16-byte aligned and padded); free with device_free."""
        a, n, keep = _as_pointer(data)
        out = ctypes.c_void_p()
        self.lib.check(self.lib.L.eds_device_upload(self.handle, a, n, ctypes.byref(out)))
        del keep
        return Buffer(out.value, n)

    def device_free(self, buf):
        self.lib.L.eds_device_free(self.handle, buf.data)

    def vcf_transform_device(self, vcf_buf, fasta_buf):
        """device Buffers in -> (eds Buffer, seds Buffer, stats dict), outputs stay on the device (l = 0 only)."""
        e, s, st = Buffer(), Buffer(), VcfStats()
        self.lib.check(self.lib.L.eds_vcf_transform_device(self.handle, vcf_buf.data, vcf_buf.bytes, fasta_buf.data,
                                                           fasta_buf.bytes, ctypes.byref(e), ctypes.byref(s), ctypes.byref(st)))
        return e, s, st.as_dict()


_product = None


def load():
    """The product library (nvcc build). Raises ImportError when it has not been built."""
    global _product
    if _product is None:
        _product = Library(PRODUCT_SO)
    return _product
