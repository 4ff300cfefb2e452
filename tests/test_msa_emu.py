"""CPU tier: the MSA kernels' logic, stepped through the test-only CUDA emulator build, against the
oracle and the reference's golden vectors. (Parity proper is tests/test_msa_gpu.py on a B200.)"""
import pytest

import emu_lib
import msa_checks


@pytest.fixture(scope="module")
def ctx():
    c = emu_lib.lib().context()
    c.set_tuning(3, 1)  # few partitions: every scan crosses partition edges even on tiny inputs
    yield c
    c.close()


def test_reference_unit_strings(ctx):
    msa_checks.check_reference_unit_strings(ctx)


def test_golden_subset(ctx):
    assert msa_checks.check_golden(ctx, stride=16) >= 35


def test_random_against_oracle(ctx):
    msa_checks.check_random_against_oracle(ctx, seed=1, n_cases=40, max_cols=120)


def test_leds_flag_with_l0(ctx):
    msa_checks.check_leds_flag_with_l0(ctx)


def test_conserved_bits(ctx):
    msa_checks.check_conserved_bits(ctx, on_gpu=False, n_cases=4)


def test_bad_inputs(ctx):
    msa_checks.check_bad_inputs(ctx)


def test_hash_collision_fallback():
    msa_checks.check_hash_collision_fallback(emu_lib.lib(), n_cases=6)


def test_wide_alphabet(ctx):
    msa_checks.check_wide_alphabet(ctx, n_cases=5)


def test_narrow_path_off():
    msa_checks.check_narrow_off(emu_lib.lib(), n_cases=6, wide_cases=2)


def test_group_block_per_symbol():
    msa_checks.check_group_cta(emu_lib.lib(), n_cases=4, max_rows=50)


def test_row_sliced_scan():
    msa_checks.check_row_slices(emu_lib.lib(), n_cases=6)


def test_shards(ctx):
    msa_checks.check_shards(ctx, on_gpu=False, seed=2, n_cases=12, max_cols=150)


def test_synth_small(ctx):
    msa_checks.check_synth(ctx, n_rows=5, n_cols=900, wrap=80, l=10, variable_ppm=40000, shards=1)


def test_partition_count_does_not_matter():
    c = emu_lib.lib().context()
    try:
        for parts in (1, 2, 7):
            c.set_tuning(parts, 1)
            msa_checks.check_random_against_oracle(c, seed=9, n_cases=5, max_cols=150)
    finally:
        c.close()


def test_fused_scan_path():
    """k_scan_l2 (the default) and k_scan_fused (EDSB_FUSED_L2=0; one CTA per tile under the emulator: clusters are
    GPU-only) against the oracle."""
    import os

    for l2, name in (("1", "k_scan_l2"), ("0", "k_scan_fused")):
        os.environ.update({"EDSB_FUSED_MIN_ROWS": "2", "EDSB_FUSED_L2": l2})
        try:
            c = emu_lib.lib().context()
        finally:
            del os.environ["EDSB_FUSED_MIN_ROWS"], os.environ["EDSB_FUSED_L2"]
        try:
            c.set_tuning(3, 1)
            c.set_profiling(True)
            msa_checks.check_random_against_oracle(c, seed=3, n_cases=12, max_cols=700)
            assert name in [n for n, _ in c.kernel_times()]
            msa_checks.check_shards(c, on_gpu=False, seed=4, n_cases=4, max_cols=150)
            msa_checks.check_synth(c, n_rows=40, n_cols=2500, wrap=80, l=10, variable_ppm=40000, shards=1)
            msa_checks.check_synth(c, n_rows=150, n_cols=3000, wrap=60, l=3, variable_ppm=30000, shards=1)  # more than 128 rows
        finally:
            c.close()
    # rows that bypass the ring (EDSB_FUSED_DIRECT): a few, and as many as there can be (half of a CTA's rows at most)
    for direct, rows in (("3", 20), ("32", 90), ("0", 30)):
        os.environ.update({"EDSB_FUSED_MIN_ROWS": "2", "EDSB_FUSED_DIRECT": direct, "EDSB_FUSED_STAGES": "2", "EDSB_FUSED_SPLIT": "1" if direct == "0" else "0"})
        try:
            c = emu_lib.lib().context()
        finally:
            for k in ("EDSB_FUSED_MIN_ROWS", "EDSB_FUSED_DIRECT", "EDSB_FUSED_STAGES", "EDSB_FUSED_SPLIT"):
                del os.environ[k]
        try:
            c.set_tuning(3, 1)
            msa_checks.check_synth(c, n_rows=rows, n_cols=3000, wrap=70, l=10, variable_ppm=30000, shards=2)
            msa_checks.check_random_against_oracle(c, seed=6, n_cases=4, max_cols=700)
        finally:
            c.close()
    # stage reuse (few stages, many tiles per CTA), tiles fetched in pairs (EDSB_FUSED_PAIR=1) and one by one, odd and even tile counts
    for stages, pair, n_cols in (("2", "1", 7000), ("4", "1", 6300), ("2", "0", 4000)):
        os.environ.update({"EDSB_FUSED_MIN_ROWS": "2", "EDSB_FUSED_STAGES": stages, "EDSB_FUSED_PAIR": pair, "EDSB_FUSED_L2": "0"})
        try:
            c = emu_lib.lib().context()
        finally:
            for k in ("EDSB_FUSED_MIN_ROWS", "EDSB_FUSED_STAGES", "EDSB_FUSED_PAIR", "EDSB_FUSED_L2"):
                del os.environ[k]
        try:
            c.set_tuning(3, 1)
            msa_checks.check_synth(c, n_rows=34, n_cols=n_cols, wrap=70, l=10, variable_ppm=30000, shards=1)
        finally:
            c.close()
