#!/usr/bin/env python
"""A few msa2eds -l 10 transforms of a resident synthetic alignment, for ncu: python tools/profile_msa.py ROWS COLS [REPS]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import edsparser_b200 as E  # noqa: E402

rows, cols = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lib = E.load()
c = lib.context(0)
view = c.msa_synth(rows, cols, 80, seed=1, variable_ppm=10_000)
for _ in range(reps):
    e, s, st = c.msa_transform_device(view, 10)
print("ok", st)
c.close()
