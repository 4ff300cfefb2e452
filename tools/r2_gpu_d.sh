#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_msa_gpu.py -x -q -m gpu > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
timeout 150 python tools/sweep_fused.py 100 10000000 > gpurun_out/r2d_sweep_c2.jsonl 2> gpurun_out/r2d_sweep_c2.err
timeout 150 python tools/sweep_fused.py 1000 3000000 > gpurun_out/r2d_sweep_c4.jsonl 2> gpurun_out/r2d_sweep_c4.err
tail -3 gpurun_out/r2d_pytest.log
