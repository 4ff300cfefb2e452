"""CPU tier: the l-EDS merge kernels' logic through the test-only CUDA emulator build."""
import pytest

import emu_lib
import leds_checks


@pytest.fixture(scope="module")
def ctx():
    c = emu_lib.lib().context()
    c.set_tuning(3, 1)
    yield c
    c.close()


def test_golden_subset(ctx):
    n, n_err = leds_checks.check_golden(ctx, stride=10)  # (the GPU tier runs all of them)
    assert n >= 48 and n_err >= 2


def test_survey_vectors(ctx):
    leds_checks.check_survey_vectors(ctx)


def test_errors(ctx):
    leds_checks.check_errors(ctx)


def test_random(ctx):
    leds_checks.check_random(ctx, seed=3, n_cases=40, max_sym=14)


def test_long_strings_and_many_paths(ctx):
    leds_checks.check_long_strings_and_many_paths(ctx)


def test_msa_pipeline(ctx):
    leds_checks.check_msa_pipeline(ctx, n_cases=2)


def test_genrandomeds_device_matches_numpy(ctx):
    """eds_genrandomeds_device (genrandomeds-shaped EDS + SEDS generated on the device) == its numpy statement, and the
    merge of that device-resident text == the oracle."""
    import oracle_lib
    from edsparser_b200 import synth

    for n, ppm, paths, seed in ((400, 100_000, 4, 1), (300, 300_000, 3, 2), (250, 0, 4, 5)):
        e, s = ctx.genrandomeds_device(n, ppm, paths, seed)
        got = (ctx.download(e), ctx.download(s))
        assert got == synth.genrandomeds(n, ppm, paths, seed)
        assert ctx.leds_merge_device_in(e, s, 3)[:2] == oracle_lib.eds2leds(got[0], got[1], 3)
