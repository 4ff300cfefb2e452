"""GPU tier (-m gpu): byte parity of the sm_100a MSA path with the oracle, through the C ABI."""
import pytest

import edsparser_b200
import msa_checks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return edsparser_b200.load()  # raises when the CUDA library is not built: no CPU fallback


@pytest.fixture(scope="module")
def ctx(lib):
    c = lib.context(0)
    yield c
    c.close()


def test_reference_unit_strings(ctx):
    msa_checks.check_reference_unit_strings(ctx)


def test_golden_all(ctx):
    assert msa_checks.check_golden(ctx) > 500


def test_random_against_oracle(ctx):
    msa_checks.check_random_against_oracle(ctx, seed=1, n_cases=300, max_cols=400)
    msa_checks.check_random_against_oracle(ctx, seed=2, n_cases=40, max_rows=150, max_cols=3000)


def test_leds_flag_with_l0(ctx):
    msa_checks.check_leds_flag_with_l0(ctx)


def test_conserved_bits(ctx):
    msa_checks.check_conserved_bits(ctx, on_gpu=True, n_cases=20)


def test_bad_inputs(ctx):
    msa_checks.check_bad_inputs(ctx)


def test_hash_collision_fallback(lib):
    msa_checks.check_hash_collision_fallback(lib, n_cases=60)


def test_wide_alphabet(ctx):
    msa_checks.check_wide_alphabet(ctx, n_cases=30)


def test_narrow_path_off(lib):
    msa_checks.check_narrow_off(lib, n_cases=60)


def test_row_sliced_scan(lib):
    msa_checks.check_row_slices(lib, n_cases=60)


def test_shards(ctx):
    msa_checks.check_shards(ctx, on_gpu=True, seed=2, n_cases=80, max_cols=2000)


def test_partition_count_does_not_matter(lib):
    c = lib.context(0)
    try:
        for parts in (1, 3, 64, 1024):
            c.set_tuning(parts, 0)
            msa_checks.check_random_against_oracle(c, seed=9, n_cases=15, max_cols=2000)
    finally:
        c.close()


def test_synth_config2_shape_small(ctx):
    # BASELINE config 2 shape (100 rows, 1% variable columns, wrap 80, l = 10) at 200 kbp
    st = msa_checks.check_synth(ctx, n_rows=100, n_cols=200_000, wrap=80, l=10, shards=4)
    assert st["n_variable_cols"] > 1000


def test_synth_many_rows(ctx):
    # config 4 shape: 1000 rows (shared-memory budget of k_group changes), plain EDS and l-EDS
    msa_checks.check_synth(ctx, n_rows=1000, n_cols=30_000, wrap=80, l=10, shards=2)
    msa_checks.check_synth(ctx, n_rows=1000, n_cols=10_000, wrap=60, l=0)


def test_synth_wide_rows(ctx):
    # R > 4096: per-warp scratch moves to global memory; R > 65535 would switch ids to uint32
    msa_checks.check_synth(ctx, n_rows=5000, n_cols=3_000, wrap=70, l=5, variable_ppm=20000)


def test_synth_cluster_sizes(lib):
    # k_scan_l2 (default) takes any depth up to 2048 rows in one CTA; k_scan_fused (EDSB_FUSED_L2=0) splits the rows of a
    # tile over a thread-block cluster: 1 CTA up to 128 rows, then 2, 4, 8
    import os

    for l2 in ("1", "0"):
        os.environ["EDSB_FUSED_L2"] = l2
        try:
            c = lib.context(0)
        finally:
            del os.environ["EDSB_FUSED_L2"]
        try:
            for rows in (33, 128, 129, 200, 300, 520, 777, 1024):
                st = msa_checks.check_synth(c, n_rows=rows, n_cols=40_000, wrap=80, l=10, variable_ppm=15000, shards=2)
                assert st["n_variable_cols"] > 300
        finally:
            c.close()


def test_fused_scan_matches_two_pass(lib):
    # the same alignments through k_scan + k_stash (EDSB_FUSED=0), k_scan_fused (the TMA ring) and k_scan_l2 (the default)
    import os

    for fused, l2, name in (("0", "1", "k_scan"), ("1", "0", "k_scan_fused"), ("1", "1", "k_scan_l2")):
        os.environ.update({"EDSB_FUSED": fused, "EDSB_FUSED_L2": l2})
        try:
            c = lib.context(0)
        finally:
            del os.environ["EDSB_FUSED"], os.environ["EDSB_FUSED_L2"]
        try:
            c.set_profiling(True)
            msa_checks.check_synth(c, n_rows=100, n_cols=150_000, wrap=80, l=10)
            names = [n for n, _ in c.kernel_times()]
            assert name in names and ("k_stash" in names) == (fused == "0")
            c.set_profiling(False)
            msa_checks.check_random_against_oracle(c, seed=5, n_cases=30, max_rows=150, max_cols=3000)
        finally:
            c.close()


def test_group_block_per_symbol(lib):
    msa_checks.check_group_cta(lib, n_cases=25, max_rows=300)
    # a deep alignment with wide symbols: the few symbols the tuple form does not take get a block each
    c = lib.context(0)
    try:
        st = msa_checks.check_synth(c, n_rows=1000, n_cols=60_000, wrap=80, l=10, variable_ppm=40000, shards=1)
        assert st["n_hashed_symbols"] > 0
    finally:
        c.close()
