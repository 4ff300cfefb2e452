"""ctypes loader for the CPU restatement (oracle/libeds_oracle.so). Test infrastructure only."""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None

STATUS_EXC = {1: RuntimeError, 2: ValueError, 3: IndexError, 4: Exception}


class OracleError(Exception):
    def __init__(self, status, message):
        super().__init__(message)
        self.status = status
        self.message = message


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ORACLE_DIR, "libeds_oracle.so")
        src = os.path.join(ORACLE_DIR, "eds_oracle.cpp")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "restatement"], stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(so)
        cp, sz = ctypes.c_char_p, ctypes.c_size_t
        pp, ps = ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)
        L.oracle_msa2eds.argtypes = [cp, sz, ctypes.c_int, sz, pp, ps, pp, ps, cp, sz]
        L.oracle_msa2eds.restype = ctypes.c_int
        L.oracle_msa_conserved.argtypes = [cp, sz, ctypes.c_void_p, sz, cp, sz]
        L.oracle_msa_conserved.restype = ctypes.c_longlong
        L.oracle_eds2leds.argtypes = [cp, sz, cp, sz, ctypes.c_uint, ctypes.c_int, sz, pp, ps, pp, ps, cp, sz]
        L.oracle_eds2leds.restype = ctypes.c_int
        L.oracle_vcf2eds.argtypes = [cp, sz, cp, sz, sz, pp, ps, pp, ps, ctypes.POINTER(ctypes.c_ulonglong), pp, cp, sz]
        L.oracle_vcf2eds.restype = ctypes.c_int
        L.oracle_free.argtypes = [ctypes.c_void_p]
        _LIB = L
    return _LIB


def _take(L, ptr, n):
    data = ctypes.string_at(ptr.value, n.value) if ptr.value else b""
    L.oracle_free(ptr)
    return data


def msa2eds(text: bytes, l: int):
    """Reference semantics of the msa2eds CLI: l == 0 -> EDS, l > 0 -> l-EDS."""
    L = lib()
    e, s = ctypes.c_void_p(), ctypes.c_void_p()
    en, sn = ctypes.c_size_t(), ctypes.c_size_t()
    err = ctypes.create_string_buffer(512)
    rc = L.oracle_msa2eds(text, len(text), 1 if l > 0 else 0, l, ctypes.byref(e), ctypes.byref(en), ctypes.byref(s),
                          ctypes.byref(sn), err, 512)
    if rc:
        raise OracleError(rc, err.value.decode("latin-1"))
    return _take(L, e, en), _take(L, s, sn)


def msa_conserved(text: bytes):
    """B bit vector (one byte per column, sentinel included) of msa_transforms.cpp:36-90."""
    L = lib()
    cap = len(text) + 2
    buf = ctypes.create_string_buffer(cap)
    err = ctypes.create_string_buffer(512)
    C = L.oracle_msa_conserved(text, len(text), buf, cap, err, 512)
    if C < 0:
        raise OracleError(1, err.value.decode("latin-1"))
    return bytes(buf.raw[: C + 1])


def eds2leds(eds: bytes, seds, l: int, compact: bool = True, max_out_bytes: int = 0):
    """seds None -> eds_to_leds_cartesian, else eds_to_leds_linear. Returns (leds, seds_out or None)."""
    L = lib()
    e, s = ctypes.c_void_p(), ctypes.c_void_p()
    en, sn = ctypes.c_size_t(), ctypes.c_size_t()
    err = ctypes.create_string_buffer(512)
    rc = L.oracle_eds2leds(eds, len(eds), seds, 0 if seds is None else len(seds), l, 1 if compact else 0,
                           max_out_bytes, ctypes.byref(e), ctypes.byref(en), ctypes.byref(s), ctypes.byref(sn), err, 512)
    if rc:
        raise OracleError(rc, err.value.decode("latin-1"))
    out, sout = _take(L, e, en), _take(L, s, sn)
    return out, (sout if seds is not None else None)


VCF_STAT_KEYS = ("total", "processed", "malformed", "sv", "groups")


def vcf2eds(vcf: bytes, fasta: bytes, l: int = 0):
    """Reference semantics of the vcf2eds CLI: l == 0 -> parse_vcf_to_eds_streaming, l > 0 -> ..._to_leds_...
    Returns (eds, seds, stats dict, warning lines)."""
    L = lib()
    e, s, w = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    en, sn = ctypes.c_size_t(), ctypes.c_size_t()
    st = (ctypes.c_ulonglong * 5)()
    err = ctypes.create_string_buffer(512)
    rc = L.oracle_vcf2eds(vcf, len(vcf), fasta, len(fasta), l, ctypes.byref(e), ctypes.byref(en), ctypes.byref(s),
                          ctypes.byref(sn), st, ctypes.byref(w), err, 512)
    if rc:
        raise OracleError(rc, err.value.decode("latin-1"))
    warn = ctypes.string_at(w.value).decode("latin-1").splitlines() if w.value else []
    L.oracle_free(w)
    return _take(L, e, en), _take(L, s, sn), dict(zip(VCF_STAT_KEYS, [int(x) for x in st])), warn
