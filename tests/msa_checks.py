"""Parity checks of the MSA path, written once and run against
  * the product library on a B200 (tests/test_msa_gpu.py, -m gpu) — the parity tier proper, and
  * the g++/emulator build of the same kernel sources (tests/test_msa_emu.py, CPU tier) — kernel-logic
    checks that need no GPU.
Both go through the C ABI; the oracle (tests/oracle_lib.py) is only the checker."""
import ctypes
import json
import os

import numpy as np

import gen
import oracle_lib
from edsparser_b200 import capi, synth

HERE = os.path.dirname(os.path.abspath(__file__))


class DeviceText:
    """A byte string in 'device' memory of the library under test (cuda tensor, or host memory for
    the emulator build)."""

    def __init__(self, data: bytes, on_gpu: bool):
        self.n = len(data)
        if on_gpu:
            import torch

            self.t = torch.frombuffer(bytearray(data + b"\0" * 64), dtype=torch.uint8).cuda()
            self.ptr = self.t.data_ptr()
        else:
            self.t = np.zeros(len(data) + 80, dtype=np.uint8)
            off = (-self.t.ctypes.data) % 16
            self.t[off:off + len(data)] = np.frombuffer(data, dtype=np.uint8)
            self.ptr = self.t.ctypes.data + off
        assert self.ptr % 16 == 0


def make_view(dev, idx, col_begin=0, col_count=None, own=None):
    """eds_msa_view over a whole file held in `dev`, restricted to a column window."""
    C, lw = idx["n_cols"], idx["line_width"]
    if col_count is None:
        col_count = C - col_begin
    rows = (ctypes.c_uint64 * idx["n_rows"])(*[s + col_begin + col_begin // lw for s in idx["row_start"]])
    v = capi.MsaView()
    v.text = dev.ptr
    v.text_bytes = dev.n
    v.row_start = ctypes.cast(rows, ctypes.POINTER(ctypes.c_uint64))
    v.n_rows = idx["n_rows"]
    v.line_width = lw
    v.total_cols = C
    v.col_begin, v.col_count = col_begin, col_count
    v.own_begin, v.own_end = own if own else (col_begin, col_begin + col_count)
    v._keep = (rows, dev)
    return v


def golden_cases(stride=1):
    with open(os.path.join(HERE, "golden", "msa.json")) as f:
        return json.load(f)["cases"][::stride]


def check_golden(ctx, stride=1):
    n = 0
    for c in golden_cases(stride):
        text = c["msa"].encode("latin-1")
        e, s, _ = ctx.msa_transform_host(text, c["l"])
        assert e == c["eds"].encode("latin-1"), (c["msa"], c["l"], e)
        assert s == c["seds"].encode("latin-1"), (c["msa"], c["l"], s)
        n += 1
    return n


def check_reference_unit_strings(ctx):
    # golden strings of the reference's own tests/cpp/test_msa.cpp:20-229
    msa = b">seq1\nAGTC--TCTATA\n>seq2\nAGTCCCTATATA\n>seq3\nAGTC--TATATA\n"
    assert ctx.msa_transform_host(msa, 0)[:2] == (b"{AGTC}{,CC}{T}{C,A}{TATA}", b"{0}{1,3}{2}{0}{1}{2,3}{0}")
    assert ctx.msa_transform_host(msa, 4)[:2] == (b"{AGTC}{TC,CCTA,TA}{TATA}", b"{0}{1}{2}{3}{0}")
    same = b">a\nACGTACGT\n>b\nACGTACGT\n>c\nACGTACGT\n"
    assert ctx.msa_transform_host(same, 0)[:2] == (b"{ACGTACGT}", b"{0}")
    snp = b">a\nACGT\n>b\nACGA\n"
    assert ctx.msa_transform_host(snp, 0)[:2] == (b"{ACG}{T,A}", b"{0}{1}{2}")
    # SURVEY.md Appendix B (data/msa/small.msa, wrapped at 12)
    small = (b">seq1\nAGTC--TCTATA\nAATAAATA----\n>seq2\nAGTCCCTATATA\nAATAAATAGGGG\n"
             b">seq3\nAGTC--TATATA\nAATGGATA----\n")
    assert ctx.msa_transform_host(small, 0)[:2] == (
        b"{AGTC}{,CC}{T}{C,A}{TATAAAT}{AA,GG}{ATA}{,GGGG}", b"{0}{1,3}{2}{0}{1}{2,3}{0}{1,2}{3}{0}{1,3}{2}")
    assert ctx.msa_transform_host(small, 10)[:2] == (
        b"{AGTC}{TCTATAAATAAATA,CCTATATAAATAAATAGGGG,TATATAAATGGATA}", b"{0}{1}{2}{3}")
    # gaps only: every row '-' in a column is NOT conserved (msa_transforms.cpp:75)
    assert ctx.msa_transform_host(b">a\nA-GT\n>b\nA-GT\n", 0)[:2] == (b"{A}{}{GT}", b"{0}{1,2}{0}")
    assert ctx.msa_transform_host(b">a\nacgt\n>b\nACGT\n", 0)[:2] == (b"{acgt,ACGT}", b"{1}{2}")


def check_random_against_oracle(ctx, seed, n_cases, max_rows=9, max_cols=200, ls=(0, 1, 2, 3, 5, 10, 1000)):
    rng = np.random.default_rng(seed)
    for i in range(n_cases):
        text, m, wrap = gen.random_msa_text(rng, max_rows=max_rows, max_cols=max_cols)
        l = int(rng.choice(ls))
        exp = oracle_lib.msa2eds(text, l)
        got = ctx.msa_transform_host(text, l)
        if i % 4 == 0:
            assert ctx.msa_transform_host_view(text, l)[:2] == got[:2]  # pinned-view form of the same call
        assert got[0] == exp[0], (seed, i, l, wrap, text)
        assert got[1] == exp[1], (seed, i, l, wrap, text)


def check_leds_flag_with_l0(ctx):
    # parse_msa_to_leds_streaming(in, 0) through the API (the CLI never does this): every conserved
    # run is standalone, so it equals the plain EDS.
    rng = np.random.default_rng(5)
    text, _, _ = gen.random_msa_text(rng, max_cols=120)
    a = ctx.msa_transform_host(text, 0, leds=False)[:2]
    b = ctx.msa_transform_host(text, 0, leds=True)[:2]
    assert a == b == oracle_lib.msa2eds(text, 0)


def check_conserved_bits(ctx, on_gpu, seed=3, n_cases=6):
    rng = np.random.default_rng(seed)
    for _ in range(n_cases):
        text, m, wrap = gen.random_msa_text(rng, max_cols=300)
        idx = ctx.msa_index(text)
        dev = DeviceText(text, on_gpu)
        bits = np.unpackbits(np.frombuffer(ctx.msa_conserved_bits(make_view(dev, idx)), dtype=np.uint8),
                             bitorder="little")[: idx["n_cols"]]
        exp = np.frombuffer(oracle_lib.msa_conserved(text), dtype=np.uint8)[: idx["n_cols"]]
        assert (bits == exp).all(), (text,)
        # numpy statement of msa_transforms.cpp:75: all rows equal row 0 and row 0 is not '-'
        assert (bits == ((m == m[0]).all(axis=0) & (m[0] != ord("-")))).all()


def check_bad_inputs(ctx):
    def status(text, l=0):
        try:
            ctx.msa_transform_host(text, l)
        except capi.EdsError as e:
            return e.status
        return 0

    assert status(b">only\nACGT\n") == capi.EDS_ERR_BAD_MSA  # R = 1 is undefined in the reference
    assert status(b"ACGT\n") == capi.EDS_ERR_BAD_MSA
    assert status(b">a\nACGT\n>b\nACG\n") == capi.EDS_ERR_BAD_MSA  # short row
    assert status(b">a\nACGT\n>b\nACGTA\n") == capi.EDS_ERR_BAD_MSA  # long row
    assert status(b">a\nACGT\nAC\n>b\nACG\nTAC\n") == capi.EDS_ERR_BAD_MSA  # different wrap
    assert status(b">a\nACGT\nAC\n>b\nACGT\n\nAC\n") == capi.EDS_ERR_BAD_MSA  # blank line inside a record
    assert status(b">a\n>b\nACGT\n") == capi.EDS_ERR_BAD_MSA  # empty first row
    # accepted: blank line before a header, missing final newline, blank lines at the end
    ok = b">a\nACGT\nAC\n\n>b\nACGA\nAC"
    assert ctx.msa_transform_host(ok, 0)[:2] == oracle_lib.msa2eds(ok, 0)
    ok = b">a\nACGT\n>b\nACGA\n\n\n"
    assert ctx.msa_transform_host(ok, 0)[:2] == oracle_lib.msa2eds(ok, 0)


def check_hash_collision_fallback(lib, seed=11, n_cases=25):
    """EDSB_DEBUG_HASH_MASK keeps 2 bits of the row hash: nearly every symbol takes the exact
    quadratic fallback, and the output must not change."""
    os.environ["EDSB_DEBUG_HASH_MASK"] = "0x3"
    try:
        ctx = lib.context()
    finally:
        del os.environ["EDSB_DEBUG_HASH_MASK"]
    try:
        check_random_against_oracle(ctx, seed, n_cases, max_rows=12, max_cols=60, ls=(0, 3, 10))
    finally:
        ctx.close()


def check_wide_alphabet(ctx, seed=21, n_cases=8):
    """More than 8 distinct residues in a column: the lane-per-symbol path must hand the symbol to the
    warp-per-symbol path (its register list holds 8)."""
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWY", dtype=np.uint8)
    for i in range(n_cases):
        m = gen.random_alignment(rng, 40, int(rng.integers(5, 80)), p_var=0.2, p_sub=0.8, p_gap=0.05, alphabet=alphabet)
        text = gen.to_fasta(m, 60)
        l = int(rng.choice([0, 3]))
        assert ctx.msa_transform_host(text, l)[:2] == oracle_lib.msa2eds(text, l), (seed, i, l, text)


def check_narrow_off(lib, seed=12, n_cases=15, wide_cases=3, max_rows=70, ls=(0, 2, 10)):
    """EDSB_DEBUG_NARROW_OFF=1: single-column symbols take the rows-across-lanes path (what large R uses);
    =2: every variable symbol goes through the hashed warp-per-symbol path."""
    for mode in ("1", "2"):
        os.environ["EDSB_DEBUG_NARROW_OFF"] = mode
        try:
            ctx = lib.context()
        finally:
            del os.environ["EDSB_DEBUG_NARROW_OFF"]
        try:
            check_random_against_oracle(ctx, seed, n_cases, max_rows=max_rows, max_cols=100, ls=ls)
            check_wide_alphabet(ctx, n_cases=wide_cases)
        finally:
            ctx.close()


def check_group_cta(lib, seed=21, n_cases=8, max_rows=70):
    """EDSB_DEBUG_GROUP_CTA=2 with EDSB_DEBUG_NARROW_OFF=2: every variable symbol goes through the hashed path with a whole
    block per symbol (what the few wide symbols of a deep alignment get); =0: a warp per symbol."""
    for mode in ("2", "0"):
        os.environ.update({"EDSB_DEBUG_GROUP_CTA": mode, "EDSB_DEBUG_NARROW_OFF": "2"})
        try:
            ctx = lib.context()
        finally:
            del os.environ["EDSB_DEBUG_GROUP_CTA"], os.environ["EDSB_DEBUG_NARROW_OFF"]
        try:
            check_random_against_oracle(ctx, seed, n_cases, max_rows=max_rows, max_cols=120, ls=(0, 3, 10))
            check_wide_alphabet(ctx, n_cases=2)
        finally:
            ctx.close()
    # forced hash collisions (exact fallback search) through the block form
    os.environ.update({"EDSB_DEBUG_GROUP_CTA": "2", "EDSB_DEBUG_NARROW_OFF": "2", "EDSB_DEBUG_HASH_MASK": "0x3"})
    try:
        ctx = lib.context()
    finally:
        del os.environ["EDSB_DEBUG_GROUP_CTA"], os.environ["EDSB_DEBUG_NARROW_OFF"], os.environ["EDSB_DEBUG_HASH_MASK"]
    try:
        check_random_against_oracle(ctx, seed + 1, max(2, n_cases // 2), max_rows=max_rows, max_cols=120, ls=(0, 10))
    finally:
        ctx.close()


def check_row_slices(lib, seed=14, n_cases=15, settings=("2", "5")):
    """EDSB_DEBUG_ROW_SLICES=n: k_scan (the two-pass form, EDSB_FUSED=0) splits the rows into n slices that OR their
    mismatch bits together (what a narrow column shard of an alignment too deep for the fused scan uses)."""
    for slices in settings:
        os.environ["EDSB_DEBUG_ROW_SLICES"] = slices
        os.environ["EDSB_FUSED"] = "0"
        try:
            ctx = lib.context()
        finally:
            del os.environ["EDSB_DEBUG_ROW_SLICES"]
            del os.environ["EDSB_FUSED"]
        try:
            check_random_against_oracle(ctx, seed, n_cases, max_rows=40, max_cols=120, ls=(0, 3, 10))
        finally:
            ctx.close()


def shard_concat(ctx, dev, idx, cuts, halo, l):
    """Run every shard [cuts[i], cuts[i+1]) with `halo` columns on each side; concatenated outputs."""
    C = idx["n_cols"]
    eds, seds = b"", b""
    stats = []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        wb, we = max(0, lo - halo), min(C, hi + halo)
        v = make_view(dev, idx, wb, we - wb, (lo, hi))
        e, s, st = ctx.msa_transform_device(v, l)
        eds += ctx.download(e)
        seds += ctx.download(s)
        stats.append(st)
    return eds, seds, stats


def check_shards(ctx, on_gpu, seed, n_cases, max_cols=400):
    rng = np.random.default_rng(seed)
    done = 0
    for i in range(n_cases):
        text, m, wrap = gen.random_msa_text(rng, max_rows=6, max_cols=max_cols)
        idx = ctx.msa_index(text)
        C = idx["n_cols"]
        if C < 8:
            continue
        l = int(rng.choice([0, 1, 3, 10]))
        whole = oracle_lib.msa2eds(text, l)
        dev = DeviceText(text, on_gpu)
        n_shards = int(rng.integers(2, 6))
        inner = sorted(set(int(x) for x in rng.integers(1, C, n_shards - 1)))
        cuts = [0] + inner + [C]
        halo = max(l + 1, 1)
        while True:
            try:
                eds, seds, stats = shard_concat(ctx, dev, idx, cuts, halo, l)
                break
            except capi.EdsError as e:
                assert e.status == capi.EDS_ERR_HALO, e
                assert halo < C, "halo failure with the whole alignment in view"
                halo = min(C, halo * 2 + 1)
        assert eds == whole[0], (seed, i, l, cuts, halo, text)
        assert seds == whole[1], (seed, i, l, cuts, halo, text)
        assert sum(s["eds_bytes"] for s in stats) == len(whole[0])
        done += 1
    assert done > 0


def check_synth(ctx, n_rows, n_cols, wrap, l, seed=1, variable_ppm=10000, shards=1):
    """Synthetic alignment generated in device memory (configs 2/4 shape): transform it there and
    compare with the oracle run on the downloaded text."""
    v = ctx.msa_synth(n_rows, n_cols, wrap, seed=seed, variable_ppm=variable_ppm)
    text = ctx.download(capi.Buffer(v.text, v.text_bytes))
    # the numpy statement of the generator (used by bench.py's CPU legs) writes the same bytes
    assert text == synth.fasta_window(n_rows, n_cols, wrap, seed=seed, variable_ppm=variable_ppm)
    e, s, st = ctx.msa_transform_device(v, l)
    got = (ctx.download(e), ctx.download(s))
    exp = oracle_lib.msa2eds(text, l)
    assert got[0] == exp[0]
    assert got[1] == exp[1]
    if shards > 1:
        # every rank generates only its own window (+halo) of the same alignment
        halo = 4096
        eds = seds = b""
        for k in range(shards):
            lo, hi = n_cols * k // shards, n_cols * (k + 1) // shards
            wb, we = max(0, lo - halo), min(n_cols, hi + halo)
            w = ctx.msa_synth(n_rows, n_cols, wrap, col_begin=wb, col_count=we - wb, seed=seed,
                              variable_ppm=variable_ppm)
            w.own_begin, w.own_end = lo, hi
            e, s, _ = ctx.msa_transform_device(w, l)
            eds += ctx.download(e)
            seds += ctx.download(s)
        assert eds == exp[0]
        assert seds == exp[1]
    ctx.msa_synth_free()
    return st
