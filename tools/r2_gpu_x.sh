#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
run() {
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 30 --warmup 3 --no-config4 > gpurun_out/r2x_$tag.json 2> gpurun_out/r2x_$tag.err
  python - $tag <<'PY'
import json,sys
try:
    l=[x for x in open('gpurun_out/r2x_%s.json' % sys.argv[1]) if x.startswith('{')][-1]
    d=json.loads(l)
    print(sys.argv[1], 'N', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'value', d['value'])
except Exception as e:
    print(sys.argv[1], 'failed', e)
PY
}
run head56 A=1
run head0 EDSB_PEER_HEADROOM=0
run head56b A=1
run head0b EDSB_PEER_HEADROOM=0
