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
