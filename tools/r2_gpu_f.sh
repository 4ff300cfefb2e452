#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/sw.py <<'PY'
import sys, json, os
sys.path.insert(0, 'tools'); sys.path.insert(0, '.')
import sweep_fused as S, edsparser_b200 as E
rows, cols = int(sys.argv[1]), int(sys.argv[2])
lib = E.load(); base = lib.context(0)
view = base.msa_synth(rows, cols, 80, seed=1, variable_ppm=10_000)
for mode, pw, pwb, pct in ((0,4,0,0),(2,4,2,50),(2,6,2,50),(2,6,3,50),(2,8,4,50),(2,6,2,35),(2,6,3,65),(2,8,3,40)):
    env = {"EDSB_FUSED": 1, "EDSB_FUSED_MODE": mode, "EDSB_FUSED_PW": pw, "EDSB_FUSED_PWB": pwb, "EDSB_FUSED_BULK_PCT": pct, "EDSB_FUSED_DW": 4}
    try:
        k = S.run(lib, view, rows, env)
        print(json.dumps({**env, "scan_ms": round(k.get("k_scan_fused", 0), 4), "sum": round(sum(k.values()), 4)}), flush=True)
    except Exception as e:
        print(json.dumps({**env, "error": str(e)[:100]}), flush=True)
PY
timeout 150 python -m pytest tests/test_msa_gpu.py -x -q -m gpu 2>&1 | tail -2
EDSB_FUSED_MODE=2 timeout 150 python -m pytest tests/test_msa_gpu.py -x -q -m gpu -k "synth or fused" 2>&1 | tail -2
timeout 120 python /tmp/sw.py 100 10000000 > gpurun_out/r2f_c2.jsonl 2>&1
timeout 120 python /tmp/sw.py 1000 3000000 > gpurun_out/r2f_c4.jsonl 2>&1
cat gpurun_out/r2f_c2.jsonl gpurun_out/r2f_c4.jsonl
