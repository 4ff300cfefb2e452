"""GPU tier: VCF front end through the C ABI of the product library, byte for byte against the reference's golden
outputs (tests/golden/vcf.json) and the oracle port."""
import random

import pytest

import oracle_lib
import vcf_checks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import edsparser_b200

    c = edsparser_b200.load().context(0)
    yield c
    c.close()


def test_shipped(ctx):
    vcf_checks.check_shipped(ctx)


def test_golden_all(ctx):
    n, n_err = vcf_checks.check_golden(ctx)
    assert n > 300 and n_err >= 10


def test_edges(ctx):
    vcf_checks.check_edges(ctx)


def test_random(ctx):
    vcf_checks.check_random(ctx, seed=11, n_cases=120, ls=(0, 2, 10))


def test_wide_matrix(ctx):
    vcf_checks.check_wide(ctx, n_samples=300, n_sites=40)
    vcf_checks.check_wide(ctx, n_samples=2504, n_sites=60, seed=6)
    vcf_checks.check_wide(ctx, n_samples=11000, n_sites=8, seed=7)   # five-digit ids


def test_many_alleles_fall_back_to_global_bitsets(ctx):
    # (alleles + 1) x words beyond the warp's shared-memory slice
    rng = random.Random(3)
    ref = "".join(rng.choice("ACGT") for _ in range(200))
    fa = (">r\n" + ref + "\n").encode()
    n_s = 2100
    alts = ",".join(ref[49] + "A" * (i + 1) for i in range(40))
    gts = "\t".join("%d|%d" % (rng.randint(0, 40), rng.randint(0, 41)) for _ in range(n_s))
    vcf = ("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\n" + "r\t50\t.\t%s\t%s\t.\t.\t.\tGT\t%s\n" % (ref[49], alts, gts)).encode()
    vcf_checks.check_case(ctx, vcf, fa, 0, None)


def test_device_resident_matches_host(ctx):
    vcf, fa = vcf_checks.synth_vcf(n_bases=20000, n_sites=300, n_samples=500, seed=2)
    exp = oracle_lib.vcf2eds(vcf, fa, 0)
    dv, df = ctx.upload(vcf), ctx.upload(fa)
    try:
        e, s, st = ctx.vcf_transform_device(dv, df)
        assert (ctx.download(e), ctx.download(s)) == exp[:2]
        assert st["variant_groups"] == exp[2]["groups"]
    finally:
        ctx.device_free(dv)
        ctx.device_free(df)
    assert ctx.vcf_transform_host(vcf, fa, 0)[:2] == exp[:2]
    assert ctx.vcf_transform_host(vcf, fa, 10)[:2] == oracle_lib.vcf2eds(vcf, fa, 10)[:2]
