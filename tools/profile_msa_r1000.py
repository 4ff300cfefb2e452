"""Small fixed config-4-shaped workload for ncu captures: 1000 sequences x 2 Mbp (2 GB), msa2eds -l 10, three passes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import edsparser_b200 as E  # noqa: E402

ctx = E.load().context(0)
v = ctx.msa_synth(1000, int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000, 80, seed=1, variable_ppm=10_000)
for _ in range(3):
    e, s, st = ctx.msa_transform_device(v, 10)
print(st)
ctx.close()
