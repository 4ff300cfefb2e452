"""CPU tier: the VCF front-end kernels' logic through the test-only CUDA emulator build."""
import pytest

import emu_lib
import vcf_checks


@pytest.fixture(scope="module")
def ctx():
    c = emu_lib.lib().context()
    c.set_tuning(3, 1)
    yield c
    c.close()


def test_shipped(ctx):
    vcf_checks.check_shipped(ctx)


def test_edges(ctx):
    vcf_checks.check_edges(ctx)


def test_golden_subset(ctx):
    n, n_err = vcf_checks.check_golden(ctx, stride=13, offset=4, max_vcf_bytes=6000)
    assert n >= 12


def test_wide(ctx):
    vcf_checks.check_wide(ctx, n_samples=70, n_sites=5)
