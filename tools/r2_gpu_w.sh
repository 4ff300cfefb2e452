#!/bin/bash
# N GPUs: group tests (msa, leds, vcf), the torchrun bench, the group vcf bench
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_group.py tests/test_cli.py -q -m gpu > gpurun_out/r2w_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log
tail -4 gpurun_out/r2w_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2w_bench_n$N.json 2> gpurun_out/r2w_bench_n$N.err
echo "bench N=$N rc=$?"
python - $N <<'PY'
import json,sys
l=[x for x in open('gpurun_out/r2w_bench_n%s.json' % sys.argv[1]) if x.startswith('{')][-1]
d=json.loads(l)
print('N', d['n_gpus'], 'ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'config4', d.get('config4',{}).get('ms_per_step'))
PY
EDSB_VCF_GPUS=$N timeout 900 python tools/bench_vcf.py 100000 > gpurun_out/r2w_vcf_group_n$N.jsonl 2> gpurun_out/r2w_vcf_group.err
python - $N <<'PY'
import json,sys
for l in open('gpurun_out/r2w_vcf_group_n%s.jsonl' % sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print({k:v for k,v in d.items() if k.startswith('group') or k.startswith('host_to_host_view')})
PY
