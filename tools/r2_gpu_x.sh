#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/r2x_final_n$N.json 2> gpurun_out/r2x_final_n$N.err
echo "bench rc=$?"
python - $N <<'PY'
import json,sys
l=[x for x in open('gpurun_out/r2x_final_n%s.json' % sys.argv[1]) if x.startswith('{')][-1]
d=json.loads(l)
print('N', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e'].get('ms_per_step'), 'config4', d.get('config4',{}).get('ms_per_step'), d.get('config4',{}).get('roofline',{}).get('step',{}).get('frac'))
PY
