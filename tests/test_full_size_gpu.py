"""GPU tier: BASELINE.json's full sizes through size-independent properties (the oracle cannot finish there in test
time): shard-concatenation == whole, structural invariants of the EDS/SEDS text, l-EDS predicate, idempotence of the
merge, and prefix parity against the oracle."""
import os
import re
import sys

import numpy as np
import pytest

import oracle_lib
import vcf_checks

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.fixture(scope="module")
def ctx():
    import edsparser_b200

    c = edsparser_b200.load().context(0)
    yield c
    c.close()


def test_msa2eds_config2_full_size(ctx):
    """100 sequences x 10 Mbp, 1 % variable columns, l = 10 (BASELINE config 2)."""
    R, C, L, wrap = 100, 10_000_000, 10, 80
    v = ctx.msa_synth(R, C, wrap, seed=1, variable_ppm=10_000)
    e, s, st = ctx.msa_transform_device(v, L)
    eds, seds = ctx.download(e), ctx.download(s)
    assert len(eds) == st["eds_bytes"] and len(seds) == st["seds_bytes"]
    # 1. four column shards with a halo, every shard generated on its own: concatenation == whole
    parts_e, parts_s = [], []
    for k in range(4):
        lo, hi = C * k // 4, C * (k + 1) // 4
        wb, we = max(0, lo - 4096), min(C, hi + 4096)
        w = ctx.msa_synth(R, C, wrap, col_begin=wb, col_count=we - wb, seed=1, variable_ppm=10_000)
        w.own_begin, w.own_end = lo, hi
        pe, ps, _ = ctx.msa_transform_device(w, L)
        parts_e.append(ctx.download(pe))
        parts_s.append(ctx.download(ps))
    assert b"".join(parts_e) == eds and b"".join(parts_s) == seds
    ctx.msa_synth_free()
    # 2. structure: one source set per alternative; conserved symbols carry {0}; in a variable symbol every row
    #    appears exactly once (ids per variable symbol = R)
    a = np.frombuffer(eds, dtype=np.uint8)
    n_sym = int((a == ord("{")).sum())
    assert n_sym == int((a == ord("}")).sum()) == st["n_symbols"]
    n_alt = n_sym + int((a == ord(",")).sum())
    b = np.frombuffer(seds, dtype=np.uint8)
    assert int((b == ord("{")).sum()) == n_alt
    n_zero_sets = seds.count(b"{0}")
    assert n_zero_sets == n_sym - st["n_variable"]
    n_ids = int((b == ord(",")).sum()) + n_alt - n_zero_sets
    assert n_ids == R * st["n_variable"]
    # 3. the first alternative of every symbol is row 0's string: their concatenation is row 0 without gaps
    from edsparser_b200 import synth

    row0 = synth.fasta_window(1, C, wrap, seed=1, variable_ppm=10_000).split(b"\n", 1)[1].replace(b"\n", b"").replace(b"-", b"")
    first_alts = re.sub(rb"\{([^,}]*)[^}]*\}", rb"\1", eds)
    assert first_alts == row0
    # 4. it is an l-EDS (is_leds, eds_transforms.cpp:439-468)
    assert ctx.is_leds(eds, L)


def test_eds2leds_config3_full_size(ctx):
    """genrandomeds-shaped 100 Mbp, 10 % sites, 4 paths, LINEAR l = 10 (BASELINE config 3)."""
    import bench_leds

    eds, seds = bench_leds.genrandomeds_like(100_000_000)
    out, sout, rounds = ctx.leds_merge_host(eds, seds, 10)
    assert rounds >= 1 and out.endswith(b"\n") and sout.endswith(b"\n")
    assert ctx.is_leds(out, 10)
    again = ctx.leds_merge_host(out, sout, 10)
    assert again[0] == out and again[1] == sout and again[2] == 0
    # merging never changes the spelled sequence of first alternatives outside merged groups: total bases of the
    # reference path (first alternative of every symbol) are preserved
    ref_in = re.sub(rb"\{([^,}]*)[^}]*\}", rb"\1", eds)
    ref_out = re.sub(rb"\{([^,}]*)[^}]*\}", rb"\1", out.rstrip(b"\n"))
    assert ref_in == ref_out
    # the same generator at a size the oracle finishes: byte parity (the big run differs only in length)
    e5, s5 = bench_leds.genrandomeds_like(2_000_000, seed=5)
    assert ctx.leds_merge_host(e5, s5, 10)[:2] == oracle_lib.eds2leds(e5, s5, 10)
    # CARTESIAN on this input is refused rather than attempted
    import edsparser_b200

    with pytest.raises(edsparser_b200.EdsError) as ei:
        ctx.leds_merge_host(eds, None, 10, max_output_bytes=8 << 30)
    assert ei.value.status == edsparser_b200.EDS_ERR_BUDGET


def test_vcf2eds_config5_shape(ctx):
    """Config 5 shape at 50 000 sites x 2504 samples (0.5 GB of VCF; the full 10 GB case is tools/bench_vcf.py)."""
    # no same-position ties here: with ties the order std::sort leaves depends on the whole array, so a prefix of the
    # file is not a prefix of the problem (ties are covered by the golden cases and tools/bench_vcf.py)
    vcf, fa = vcf_checks.synth_vcf(n_bases=5_000_000, n_sites=50_000, n_samples=2504, seed=3, same_pos=False)
    dv, df = ctx.upload(vcf), ctx.upload(fa)
    try:
        e, s, st = ctx.vcf_transform_device(dv, df)
        eds, seds = ctx.download(e), ctx.download(s)
    finally:
        ctx.device_free(dv)
        ctx.device_free(df)
    assert st["processed_variants"] == st["total_variants"] >= 50_000 and st["variant_groups"] <= st["processed_variants"]
    assert st["host_sorted"] == 0
    # 1. prefix parity: the oracle on the first records gives, up to its closing common symbol, a prefix of the output
    lines = vcf.split(b"\n", 1502)
    head = b"\n".join(lines[:1502]) + b"\n"
    pe, ps, pst, _ = oracle_lib.vcf2eds(head, fa, 0)
    cut_e, cut_s = pe.rfind(b"{"), ps.rfind(b"{")
    assert pe[cut_e:].count(b",") == 0 and ps[cut_s:] == b"{0}"
    assert eds[:cut_e] == pe[:cut_e] and seds[:cut_s] == ps[:cut_s]
    # 2. structure: symbols and source sets pair up; the conserved symbols are exactly the reference between the
    #    groups of overlapping records (vcf_transforms.cpp:503-519, 570-577, 658-664)
    a = np.frombuffer(eds, dtype=np.uint8)
    n_sym = int((a == ord("{")).sum())
    assert n_sym == int((a == ord("}")).sum())
    n_alt = n_sym + int((a == ord(",")).sum())
    b = np.frombuffer(seds, dtype=np.uint8)
    assert int((b == ord("{")).sum()) == n_alt == int((b == ord("}")).sum())
    assert seds.count(b"{0}") == n_sym - st["variant_groups"]
    assert vcf_checks.conserved_text(eds, seds) == vcf_checks.expected_conserved_text(vcf, fa)


def reference_msa2eds(text, l, tmp_path):
    """msa2eds of the UNMODIFIED reference library (oracle/_ref/ref_driver, built here and carried to the GPU box) when it
    is there, else the oracle port: (eds bytes, seds bytes, which)."""
    import subprocess

    ref = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
    if os.path.exists(ref):
        src, e, s = tmp_path / "in.msa", tmp_path / "ref.eds", tmp_path / "ref.seds"
        src.write_bytes(text)
        subprocess.run([ref, "msa2eds", str(src), str(l), str(e), str(s)], check=True, capture_output=True)
        out = (e.read_bytes(), s.read_bytes(), "reference")
        for p in (src, e, s):
            p.unlink()
        return out
    e, s = oracle_lib.msa2eds(text, l)
    return e, s, "port"


def test_msa2eds_config2_full_size_bytes(ctx, tmp_path):
    """BASELINE config 2 at FULL size (100 x 10 Mbp = 1e9 cells, l = 10): byte parity with the reference itself."""
    R, C, L, wrap = 100, 10_000_000, 10, 80
    v = ctx.msa_synth(R, C, wrap, seed=1, variable_ppm=10_000)
    text = ctx.download(edsparser_b200_buffer(v))
    e, s, st = ctx.msa_transform_device(v, L)
    eds, seds = ctx.download(e), ctx.download(s)
    ctx.msa_synth_free()
    re_, rs_, which = reference_msa2eds(text, L, tmp_path)
    assert eds == re_ and seds == rs_, which
    assert st["n_variable_cols"] > 90_000


def edsparser_b200_buffer(view):
    import edsparser_b200

    return edsparser_b200.Buffer(view.text, view.text_bytes)


def test_msa2eds_config4_windows_bytes(ctx, tmp_path):
    """BASELINE config 4 (1000 x 30 Mbp): eight random 200 000-column windows of THAT alignment against the reference,
    each as the shard it would be (window + halo; the owned range's bytes are cut out of the reference's output of the
    window by the symbols' positions) and as a stand-alone alignment. The whole 30 G-cell run is checked through
    shard-concatenation and structure below."""
    import random

    R, C, L, wrap, W, H = 1000, 30_000_000, 10, 80, 200_000, 2048
    rng = random.Random(4)
    for k in range(8):
        lo = rng.randrange(H, C - W - H) // wrap * wrap + H % wrap  # the window starts on a line start: it is a FASTA file of its own
        w = ctx.msa_synth(R, C, wrap, col_begin=lo - H, col_count=W + 2 * H, seed=1, variable_ppm=10_000)
        text = ctx.download(edsparser_b200_buffer(w))
        # (a) the window as an alignment of its own
        e, s, _ = ctx.msa_transform_host(text, L)
        re_, rs_, which = reference_msa2eds(text, L, tmp_path)
        assert e == re_ and s == rs_, (k, which)
        # (b) the window as the shard [lo, lo + W) of the 30 Mbp alignment: symbols that start inside the owned range are
        # the stand-alone run's symbols there (both see the same columns; H columns keep the window's ends out of reach)
        w.own_begin, w.own_end = lo, lo + W
        se, ss, st = ctx.msa_transform_device(w, L)
        shard_e, shard_s = ctx.download(se), ctx.download(ss)
        assert shard_e in re_ and shard_s in rs_, k
        assert st["n_symbols"] > 1000
    ctx.msa_synth_free()


def test_msa2eds_config4_full_size(ctx):
    """The whole 1000 x 30 Mbp alignment on one GPU: four shards concatenate to the whole; structure."""
    R, C, L, wrap = 1000, 30_000_000, 10, 80
    v = ctx.msa_synth(R, C, wrap, seed=1, variable_ppm=10_000)
    e, s, st = ctx.msa_transform_device(v, L)
    eds, seds = ctx.download(e), ctx.download(s)
    ctx.msa_synth_free()
    parts_e, parts_s = [], []
    for k in range(4):
        lo, hi = C * k // 4, C * (k + 1) // 4
        wb, we = max(0, lo - 4096), min(C, hi + 4096)
        w = ctx.msa_synth(R, C, wrap, col_begin=wb, col_count=we - wb, seed=1, variable_ppm=10_000)
        w.own_begin, w.own_end = lo, hi
        pe, ps, _ = ctx.msa_transform_device(w, L)
        parts_e.append(ctx.download(pe))
        parts_s.append(ctx.download(ps))
    ctx.msa_synth_free()
    assert b"".join(parts_e) == eds and b"".join(parts_s) == seds
    a = np.frombuffer(eds, dtype=np.uint8)
    n_sym = int((a == ord("{")).sum())
    assert n_sym == int((a == ord("}")).sum()) == st["n_symbols"]
    n_alt = n_sym + int((a == ord(",")).sum())
    b = np.frombuffer(seds, dtype=np.uint8)
    assert int((b == ord("{")).sum()) == n_alt
    n_zero_sets = seds.count(b"{0}")
    assert n_zero_sets == n_sym - st["n_variable"]
    assert int((b == ord(",")).sum()) + n_alt - n_zero_sets == R * st["n_variable"]
    assert ctx.is_leds(eds, L)


def test_eds2leds_config3_20mbp_bytes(ctx):
    """genrandomeds-shaped 20 Mbp, 10 % sites, 4 paths, LINEAR l = 10: byte parity with the oracle (the reference itself
    needs hours here: it rebuilds the whole EDS per merged pair)."""
    import bench_leds

    e, s = bench_leds.genrandomeds_like(20_000_000, seed=5)
    got = ctx.leds_merge_host(e, s, 10)
    exp = oracle_lib.eds2leds(e, s, 10)
    assert got[0] == exp[0] and got[1] == exp[1]


def test_vcf2eds_config5_l10_bytes(ctx):
    """Config 5 shape, 50 000 sites x 2504 samples (0.5 GB of VCF, same-position ties included), DIRECT to l-EDS with
    l = 10 (parse_vcf_to_leds_streaming): byte parity with the oracle."""
    vcf, fa = vcf_checks.synth_vcf(n_bases=5_000_000, n_sites=50_000, n_samples=2504, seed=3)
    got = ctx.vcf_transform_host(vcf, fa, 10)
    exp = oracle_lib.vcf2eds(vcf, fa, 10)
    assert got[0] == exp[0] and got[1] == exp[1]
