#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_msa_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 60 python tools/profile_msa.py 1000 3000000 2 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "rc=$?"; tail -3 gpurun_out/r2g_bench.err
