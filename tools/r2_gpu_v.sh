#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_group.py tests/test_leds_gpu.py tests/test_cli.py -q -m gpu > gpurun_out/r2v_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
tail -6 gpurun_out/r2v_pytest.log
