#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_msa_gpu.py tests/test_leds_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "rc=$?"; tail -3 gpurun_out/r2g_bench.err
EDSB_DEBUG_NARROW_OFF=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-config4 > gpurun_out/r2g_bench_lanes.json 2> gpurun_out/r2g_bench_lanes.err
echo "rc=$?"
