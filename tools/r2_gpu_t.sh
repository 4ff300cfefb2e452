#!/bin/bash
mkdir -p gpurun_out
for v in "" cw6 cw12 cw16; do
  if [ -n "$v" ]; then export EDSB_LIBRARY=$PWD/build/variants/lib$v.so; fi
  timeout 200 python tools/sweep_fused.py 100 10000000 none > gpurun_out/r2t_cw_$v.jsonl 2> gpurun_out/r2t_cw_$v.err
  timeout 200 python tools/sweep_fused.py 1000 3000000 none >> gpurun_out/r2t_cw_$v.jsonl 2>> gpurun_out/r2t_cw_$v.err
  python - "$v" <<'PY'
import json,sys
for l in open('gpurun_out/r2t_cw_%s.jsonl' % sys.argv[1]):
    d=json.loads(l); d.pop('kernels',None); d.pop('cols',None); print(sys.argv[1] or 'cw8', d)
PY
done
