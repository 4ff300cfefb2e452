#!/bin/bash
mkdir -p gpurun_out
true
rc=0
echo "pytest rc=$rc" >> gpurun_out/r2t_pytest.log
tail -3 gpurun_out/r2t_pytest.log
if [ $rc -ne 0 ]; then grep -n "Error\|assert" gpurun_out/r2t_pytest.log | head -20; exit 1; fi
timeout 300 python tools/sweep_fused.py 100 10000000 $1 > gpurun_out/r2t_sweep_r100.jsonl 2> gpurun_out/r2t_sweep_r100.err
timeout 300 python tools/sweep_fused.py 1000 3000000 $1 > gpurun_out/r2t_sweep_r1000.jsonl 2> gpurun_out/r2t_sweep_r1000.err
python - <<'PY'
import json
for f in ('r100','r1000'):
    for l in open('gpurun_out/r2t_sweep_%s.jsonl' % f):
        d=json.loads(l); k=d.pop('kernels',{}); d.pop('cols'); print(d)
PY
