#!/bin/bash
mkdir -p gpurun_out
timeout 100 python tools/sanitize_small.py > gpurun_out/r2m_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2m_plain.log; exit 1; }
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/r2m_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2m_memcheck.log
tail -6 gpurun_out/r2m_memcheck.log
timeout 600 python tools/bench_leds.py 5000000 100000000 > gpurun_out/r2m_leds_bench.jsonl 2> gpurun_out/r2m_leds_bench.err
echo "leds rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2m_leds_bench.jsonl'):
    d=json.loads(l)
    if 'device_kernel_ms' in d: print(d['bp'], d['device_kernel_ms'], d['c_abi_pinned_view_ms'], d['n_launches'], d['top_kernels_ms'], d['parity'][:40])
PY
