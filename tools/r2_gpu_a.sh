#!/bin/bash
# round 2, GPU call A: parity of the fused scan (incl. cluster sizes), bench with and without it
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.log 2>&1
timeout 900 python -m pytest tests/test_msa_gpu.py -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench_fused.json 2> gpurun_out/r2a_bench_fused.err
echo "rc=$?" >> gpurun_out/r2a_bench_fused.err
EDSB_FUSED=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2a_bench_unfused.json 2> gpurun_out/r2a_bench_unfused.err
echo "rc=$?" >> gpurun_out/r2a_bench_unfused.err
for st in 2 3; do
EDSB_FUSED_STAGES=$st timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r2a_bench_fused_s$st.json 2> gpurun_out/r2a_bench_fused_s$st.err
done
tail -3 gpurun_out/r2a_pytest.log
cat gpurun_out/r2a_bench_fused.err | tail -5
