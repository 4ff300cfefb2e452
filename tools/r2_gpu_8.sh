#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
run() {
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 30 --warmup 3 --no-config4 > gpurun_out/r2_8_$tag.json 2> gpurun_out/r2_8_$tag.err
  python - $tag <<'PY'
import json,sys
try:
    l=[x for x in open('gpurun_out/r2_8_%s.json' % sys.argv[1]) if x.startswith('{')][-1]
    d=json.loads(l)
    print(sys.argv[1], 'N', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'value', d['value'])
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
run base A=1
run noex EDSB_BENCH_NO_EXCHANGE=1
