#!/bin/bash
# group vcf (view form) bench on N GPUs; eds2leds scan-parameter variants at config-3 size
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_group.py -q -m gpu -x > gpurun_out/r2p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -3 gpurun_out/r2p_pytest.log
EDSB_VCF_GPUS=${1:-2} timeout 900 python tools/bench_vcf.py 100000 > gpurun_out/r2p_vcf_group.jsonl 2> gpurun_out/r2p_vcf_group.err
echo "vcf bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2p_vcf_group.jsonl'):
    d=json.loads(l)
    print({k:v for k,v in d.items() if k.startswith('group') or k.startswith('host_to_host_')})
PY
for v in "" items16 items16b128 items4; do
  if [ -n "$v" ]; then export EDSB_LIBRARY=$PWD/build/variants/lib$v.so; fi
  EDSB_LEDS_CHECK_BP=0 timeout 600 python tools/bench_leds.py 100000000 > gpurun_out/r2p_leds_$v.jsonl 2> gpurun_out/r2p_leds_$v.err
  echo "leds variant '$v' rc=$?"
  python - "$v" <<'PY'
import json,sys
for l in open('gpurun_out/r2p_leds_%s.jsonl' % sys.argv[1]):
    d=json.loads(l)
    k=d.get('kernels_ms',{})
    top=sorted(k.items(), key=lambda x:-x[1])[:8]
    print(sys.argv[1], {x:d[x] for x in d if 'ms' in x and x!='kernels_ms'}, top)
PY
done
