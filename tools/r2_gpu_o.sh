#!/bin/bash
# group vcf tests on one box (slices share devices when the box has fewer), CLI, then the config-5-shaped bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_group.py tests/test_cli.py tests/test_leds_gpu.py tests/test_vcf_gpu.py -q -m gpu -x --durations=5 > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -5 gpurun_out/r2o_pytest.log
nvidia-smi -L | wc -l
EDSB_VCF_GPUS=${1:-2} timeout 900 python tools/bench_vcf.py 100000 > gpurun_out/r2o_vcf_group.jsonl 2> gpurun_out/r2o_vcf_group.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r2o_vcf_group.jsonl
