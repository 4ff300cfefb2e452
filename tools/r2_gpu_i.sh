#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 300 python -m pytest tests/test_group.py tests/test_cli.py -x -q -m gpu 2>&1 | tail -5
timeout 400 python bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err
echo "rc=$?"; tail -5 gpurun_out/r2i_bench_n2.err
