#!/bin/bash
mkdir -p gpurun_out
for p in 592 1184 2368 4096; do
  EDSB_PARTITIONS=$p EDSB_LEDS_CHECK_BP=0 timeout 300 python tools/bench_leds.py 100000000 > gpurun_out/r2y_leds_p$p.jsonl 2> gpurun_out/r2y_leds_p$p.err
  python - $p <<'PY'
import json,sys
for l in open('gpurun_out/r2y_leds_p%s.jsonl' % sys.argv[1]):
    d=json.loads(l)
    if 'device_kernel_ms' in d:
        print(sys.argv[1], d['device_kernel_ms'], d.get('c_abi_pinned_view_ms'), sorted(d.get('top_kernels_ms',{}).items(), key=lambda x:-x[1])[:6])
PY
done
