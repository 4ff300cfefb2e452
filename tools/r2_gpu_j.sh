#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_full_size_gpu.py tests/test_group.py tests/test_leds_gpu.py tests/test_msa_gpu.py tests/test_ref_cpp.py tests/test_vcf_gpu.py -q -m gpu --durations=12 -k "not config2_full_size and not config3_full_size and not config5_shape" > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -25 gpurun_out/r2j_pytest.log
