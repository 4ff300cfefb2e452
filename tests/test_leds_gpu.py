"""GPU tier (-m gpu): byte parity of the sm_100a l-EDS merge with the oracle, through the C ABI."""
import pytest

import edsparser_b200
import leds_checks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = edsparser_b200.load().context(0)
    yield c
    c.close()


def test_golden_all(ctx):
    n, n_err = leds_checks.check_golden(ctx)
    assert n > 400 and n_err >= 10


def test_survey_vectors(ctx):
    leds_checks.check_survey_vectors(ctx)


def test_errors(ctx):
    leds_checks.check_errors(ctx)


def test_random(ctx):
    assert leds_checks.check_random(ctx, seed=3, n_cases=400) >= 0
    leds_checks.check_random(ctx, seed=4, n_cases=60, max_sym=200, paths=40)


def test_long_strings_and_many_paths(ctx):
    leds_checks.check_long_strings_and_many_paths(ctx)


def test_msa_pipeline(ctx):
    leds_checks.check_msa_pipeline(ctx, n_cases=20)


def test_genrandomeds_device(ctx):
    """genrandomeds on the device: bytes == the numpy statement of the same counter-based generator;
    the device-resident pair goes through eds_leds_merge_device_in == the oracle; config-3 size generates and merges."""
    import oracle_lib
    from edsparser_b200 import synth

    for n, ppm, paths, seed in ((20_000, 100_000, 4, 1), (5_000, 400_000, 7, 2), (3_000, 10_000, 2, 3)):
        e, s = ctx.genrandomeds_device(n, ppm, paths, seed)
        got = (ctx.download(e), ctx.download(s))
        assert got == synth.genrandomeds(n, ppm, paths, seed)
        assert ctx.leds_merge_device_in(e, s, 10)[:2] == oracle_lib.eds2leds(got[0], got[1], 10)
    e, s = ctx.genrandomeds_device(100_000_000, 100_000, 4, 1)  # BASELINE config 3's size, straight into HBM
    assert e.bytes > 150_000_000 and s.bytes > 100_000_000
    out, sout, rounds = ctx.leds_merge_device_in(e, s, 10)
    assert rounds >= 1 and ctx.is_leds(out, 10)
