"""Multi-GPU entry of the C ABI (eds_group / eds_comm, csrc/shard.cu): column-sharded msa2eds with the byte-count
exchange inside the library. CPU tier: the plan, halo widening and offset logic through the emulator build (devices
played in turn, counts summed on the host). GPU tier: the same on real devices (NCCL when the box has more than one)."""
import os
import tempfile

import numpy as np
import pytest

import gen
import oracle_lib


def check_group(lib, n, n_cases, max_cols, seed=5):
    rng = np.random.default_rng(seed)
    g = lib.group(list(range(n)))
    try:
        done = 0
        for i in range(n_cases):
            text, m, wrap = gen.random_msa_text(rng, max_rows=9, max_cols=max_cols)
            for l in (0, 3, 10):
                exp = oracle_lib.msa2eds(text, l)
                try:
                    got = g.msa_transform_host(text, l, halo=8)  # small halo: the widen-and-retry path runs too
                except Exception as err:
                    if "fewer columns" in str(err):
                        continue
                    raise
                assert got[:2] == exp, (n, i, l)
                assert sum(s["eds_bytes"] for s in got[2]) == len(exp[0])
                with tempfile.TemporaryDirectory() as d:
                    tot = g.msa_transform_files(text, l, os.path.join(d, "a.eds"), os.path.join(d, "a.seds"), halo=4)
                    assert open(os.path.join(d, "a.eds"), "rb").read() == exp[0]
                    assert open(os.path.join(d, "a.seds"), "rb").read() == exp[1]
                    assert tot == (len(exp[0]), len(exp[1]))
                done += 1
        assert done > 0
    finally:
        g.close()


def test_group_emulated():
    import emu_lib

    lib = emu_lib.lib()
    check_group(lib, 3, n_cases=4, max_cols=150)


def test_comm_single_rank_emulated():
    import edsparser_b200 as E
    import emu_lib

    lib = emu_lib.lib()
    c = lib.context()
    try:
        comm = E.Comm(c, None, 0, 1)
        comm.post(10, 20)
        comm.post(30, 40)
        assert comm.offsets() == (0, 0, 30, 40)
        comm.close()
    finally:
        c.close()


@pytest.mark.gpu
def test_group_on_devices():
    import ctypes

    import edsparser_b200 as E

    lib = E.load()
    cudart = ctypes.CDLL("libcudart.so")
    n_dev = ctypes.c_int(0)
    cudart.cudaGetDeviceCount(ctypes.byref(n_dev))
    for n in sorted({1, min(2, n_dev.value), min(4, n_dev.value)}):
        check_group(lib, n, n_cases=6, max_cols=3000)


@pytest.mark.gpu
def test_group_config2_shape():
    """100 rows x 400 kbp through eds_group on every device of the box == the single-device transform."""
    import ctypes

    import edsparser_b200 as E
    from edsparser_b200 import synth

    lib = E.load()
    cudart = ctypes.CDLL("libcudart.so")
    n_dev = ctypes.c_int(0)
    cudart.cudaGetDeviceCount(ctypes.byref(n_dev))
    text = synth.fasta_window(100, 400_000, 80, seed=1, variable_ppm=10_000)
    c = lib.context(0)
    try:
        one = c.msa_transform_host(text, 10)[:2]
    finally:
        c.close()
    g = lib.group(list(range(max(1, min(8, n_dev.value)))))
    try:
        assert g.msa_transform_host(text, 10)[:2] == one
    finally:
        g.close()


def check_group_leds(lib, n, sizes, ls=(2, 3)):
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import bench_leds

    g = lib.group(list(range(n)))
    try:
        sharded = 0
        for seed, bp in enumerate(sizes, start=1):
            e, s = bench_leds.genrandomeds_like(bp, seed=seed)
            for l in ls:
                for compact in (True, False):
                    exp = oracle_lib.eds2leds(e, s, l, compact=compact)
                    got = g.leds_merge_host(e, s, l, compact=compact)
                    assert got[0] == exp[0] and got[1] == exp[1], (n, bp, l, compact)
                    sharded += got[3] > 1
        e, s = bench_leds.genrandomeds_like(sizes[0], variability=0.01, seed=9)
        assert g.leds_merge_host(e, None, 3)[0] == oracle_lib.eds2leds(e, None, 3)[0]  # CARTESIAN
        # white space inside the text, or nothing to cut at: one device, same bytes
        assert g.leds_merge_host(b"{AAAA}{A,C}\n{G}{T,G}{TTTT}", b"{0}{1,2}{2,3}{0}{1,2}{2,3}{0}", 2)[:2] == \
            oracle_lib.eds2leds(b"{AAAA}{A,C}\n{G}{T,G}{TTTT}", b"{0}{1,2}{2,3}{0}{1,2}{2,3}{0}", 2)
        return sharded
    finally:
        g.close()


def test_group_leds_emulated():
    import emu_lib

    assert check_group_leds(emu_lib.lib(), 3, sizes=(1500,), ls=(3,)) >= 1


@pytest.mark.gpu
def test_group_leds_on_devices():
    """eds2leds over symbol ranges (cuts inside long conserved symbols, verified seams) == the oracle; on a one-GPU box the
    group has one device and the call takes the single-device path."""
    import ctypes

    import edsparser_b200 as E

    lib = E.load()
    cudart = ctypes.CDLL("libcudart.so")
    n_dev = ctypes.c_int(0)
    cudart.cudaGetDeviceCount(ctypes.byref(n_dev))
    n = max(1, min(4, n_dev.value))
    sharded = check_group_leds(lib, n, sizes=(200_000, 1_000_000), ls=(3, 10))
    assert n == 1 or sharded >= 1


def check_group_vcf(lib, devices, n_random, synth_sizes, ls=(0, 3)):
    """eds_group_vcf_transform_host == the single-device transform == the oracle: random pairs (overlaps, ties, malformed
    lines: whatever cannot be sliced falls back to one device) and the config-5 generator (sliced for real)."""
    import random

    import vcf_checks

    rng = random.Random(11)
    n = len(devices)
    g = lib.group(devices)
    ctx = lib.context()
    try:
        sliced = 0
        for i in range(n_random):
            vcf, fa = vcf_checks.make_golden_vcf.random_pair(rng, n_ref=300, n_sites=40, n_samples=5)
            for l in ls:
                try:
                    exp = ctx.vcf_transform_host(vcf, fa, l)
                except Exception as err:
                    with pytest.raises(type(err)) as ei:
                        g.vcf_transform_host(vcf, fa, l)
                    assert str(ei.value) == str(err)
                    continue
                got = g.vcf_transform_host(vcf, fa, l)
                assert got[:2] == exp[:2], (i, l)
                assert g.vcf_transform_host_view(vcf, fa, l)[:2] == exp[:2]
                assert vcf_checks.stats_line(got[2]) == vcf_checks.stats_line(exp[2])
                assert got[3] == exp[3]
                sliced += got[4] > 1
        sorted_on_host = 0
        for seed, (bases, sites, samples, overlap) in enumerate(synth_sizes, start=1):
            vcf, fa = vcf_checks.synth_vcf(bases, sites, samples, seed=seed, overlap_frac=overlap)
            for l in ls:
                exp = ctx.vcf_transform_host(vcf, fa, l)
                got = g.vcf_transform_host(vcf, fa, l)
                assert got[:2] == exp[:2], (bases, l)
                assert g.vcf_transform_host_view(vcf, fa, l)[:2] == exp[:2]
                assert vcf_checks.stats_line(got[2]) == vcf_checks.stats_line(exp[2])
                assert got[2]["host_sorted"] == exp[2]["host_sorted"]
                sliced += got[4] > 1
                sorted_on_host += got[4] > 1 and got[2]["host_sorted"]
            if bases <= 4000:
                o = oracle_lib.vcf2eds(vcf, fa, 0)
                assert g.vcf_transform_host(vcf, fa, 0)[:2] == (o[0], o[1])
        assert sliced > 0 and sorted_on_host > 0  # ties went through the one sort over all slices
    finally:
        ctx.close()
        g.close()


def test_group_vcf_emulated():
    import emu_lib

    check_group_vcf(emu_lib.lib(), [0, 1], n_random=2, synth_sizes=((1500, 50, 5, 0.3),))


@pytest.mark.gpu
def test_group_vcf_on_devices():
    import ctypes

    import edsparser_b200 as E

    lib = E.load()
    cudart = ctypes.CDLL("libcudart.so")
    n_dev = ctypes.c_int(0)
    cudart.cudaGetDeviceCount(ctypes.byref(n_dev))
    # fewer devices than slices: the slices share a device, each with its own context (eds_group takes repeated ids:
    # no NCCL communicator then, the counts are exchanged on the host)
    for n in (2, 4):
        check_group_vcf(lib, [i % n_dev.value for i in range(n)], n_random=12,
                        synth_sizes=((4000, 120, 8, 0.2), (200_000, 6000, 64, 0.01), (1_000_000, 20_000, 300, 0.01)))
