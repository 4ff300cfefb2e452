#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/sweep_fused.py 100 10000000 probe > gpurun_out/r2s_sweep.jsonl 2> gpurun_out/r2s_sweep.err
echo rc=$?
python - <<'PY'
import json
for l in open('gpurun_out/r2s_sweep.jsonl'):
    d=json.loads(l); d.pop('kernels',None); print(d)
PY
