"""Small end-to-end run of both paths for compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import edsparser_b200 as E, gen, oracle_lib

ctx = E.load().context(0)
rng = np.random.default_rng(1)
for i in range(6):
    text, m, wrap = gen.random_msa_text(rng, max_rows=40, max_cols=300)
    for l in (0, 10):
        assert ctx.msa_transform_host(text, l)[:2] == oracle_lib.msa2eds(text, l)
v = ctx.msa_synth(100, 20000, 80, seed=1, variable_ppm=10000)
t = ctx.download(E.Buffer(v.text, v.text_bytes))
e, s, st = ctx.msa_transform_device(v, 10)
assert (ctx.download(e), ctx.download(s)) == oracle_lib.msa2eds(t, 10)
v = ctx.msa_synth(300, 5000, 60, seed=2, variable_ppm=20000)   # rows-across-lanes path
t = ctx.download(E.Buffer(v.text, v.text_bytes))
e, s, st = ctx.msa_transform_device(v, 3)
assert (ctx.download(e), ctx.download(s)) == oracle_lib.msa2eds(t, 3)
for i in range(6):
    eds, seds = gen.random_eds(rng, n_sym=int(rng.integers(2, 40)), with_sources=bool(i % 2))
    for l in (2, 5):
        try:
            exp = oracle_lib.eds2leds(eds, seds, l, max_out_bytes=1 << 20)
        except oracle_lib.OracleError:
            continue
        got = ctx.leds_merge_host(eds, seds, l)
        assert got[:2] == exp
fa = b">chr1\nACGTACGTACGT\n"
vcf = (b"##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\tS2\n"
       b"chr1\t5\t.\tA\tC\t.\t.\t.\tGT\t0|1\t0|0\nchr1\t5\t.\tA\tG\t.\t.\t.\tGT\t0|0\t0|1\n")
for l in (0, 3):
    assert ctx.vcf_transform_host(vcf, fa, l)[:2] == oracle_lib.vcf2eds(vcf, fa, l)[:2]
v = ctx.msa_synth(1000, 4000, 80, seed=3, variable_ppm=20000)   # clusters of 8 CTAs, tuple grouping, staged emit
t = ctx.download(E.Buffer(v.text, v.text_bytes))
e, s, st = ctx.msa_transform_device(v, 10)
assert (ctx.download(e), ctx.download(s)) == oracle_lib.msa2eds(t, 10)
print("sanitize run ok")
