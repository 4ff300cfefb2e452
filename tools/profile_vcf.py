"""Small fixed VCF workload for ncu captures: config-5-shaped input (30 000 sites x 2504 samples, ~300 MB of VCF),
uploaded once, vcf2eds run three times device-resident."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import edsparser_b200 as E  # noqa: E402
import vcf_checks  # noqa: E402

n_sites = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
vcf, fa = vcf_checks.synth_vcf(n_bases=100 * n_sites, n_sites=n_sites, n_samples=2504, seed=1)
ctx = E.load().context(0)
dv, df = ctx.upload(vcf), ctx.upload(fa)
for _ in range(3):
    e, s, st = ctx.vcf_transform_device(dv, df)
print(len(vcf), len(fa), st)
ctx.device_free(dv)
ctx.device_free(df)
ctx.close()
