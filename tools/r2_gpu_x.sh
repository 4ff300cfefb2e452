#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m pytest tests/test_group.py -q -m gpu -k "comm or group_on or config2" > gpurun_out/r2x_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2x_pytest.log
for n in 1 2 $N; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --steps 20 --warmup 3 --no-config4 > gpurun_out/r2x_bench_n$n.json 2> gpurun_out/r2x_bench_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 --no-config4 > gpurun_out/r2x_bench_n$n.json 2> gpurun_out/r2x_bench_n$n.err
  fi
  python - $n <<'PY'
import json,sys
l=[x for x in open('gpurun_out/r2x_bench_n%s.json' % sys.argv[1]) if x.startswith('{')][-1]
d=json.loads(l)
print('N', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'value', d['value'], 'e2e', d['e2e']['value'])
PY
done
