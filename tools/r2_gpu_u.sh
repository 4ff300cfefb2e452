#!/bin/bash
# full GPU tier, smoke, then the bench lines of the round (N = the GPUs of the box)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2u_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log
tail -14 gpurun_out/r2u_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2u_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2u_bench_ref.json 2> gpurun_out/r2u_bench_ref.err
echo "ref rc=$?"; tail -c 400 gpurun_out/r2u_bench_ref.json
N=$(nvidia-smi -L | wc -l); cp gpurun_out/r2u_pytest.log /dev/null 2>&1
if [ "$N" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2u_bench_n$N.json 2> gpurun_out/r2u_bench_n$N.err
  echo "bench N=$N rc=$?"; tail -c 400 gpurun_out/r2u_bench_n$N.json
  EDSB_VCF_GPUS=$N timeout 900 python tools/bench_vcf.py 100000 > gpurun_out/r2u_vcf_group.jsonl 2> gpurun_out/r2u_vcf_group.err
  echo "vcf rc=$?"
fi
