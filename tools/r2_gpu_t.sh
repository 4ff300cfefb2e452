#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_msa_gpu.py -q -m gpu -x > gpurun_out/r2t_pytest.log 2>&1
rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/r2t_pytest.log
if [ $rc -ne 0 ]; then grep -n "^E " gpurun_out/r2t_pytest.log | head; exit 1; fi
for m in 1 0; do
  EDSB_DEBUG_GROUP_CTA=$m timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2t_bench_cta$m.json 2> gpurun_out/r2t_bench_cta$m.err
  python - $m <<'PY'
import json,sys
l=[x for x in open('gpurun_out/r2t_bench_cta%s.json' % sys.argv[1]) if x.startswith('{')][-1]
d=json.loads(l); c=d['config4']; k=c['roofline']['kernels_ms_serial']
print('cta', sys.argv[1], 'config2', round(d['ms_per_step'],4), 'config4', round(c['ms_per_step'],3), {x:k[x] for x in ('k_group','k_group2','k_group3','k_emit2','k_emit_var','k_emit3')})
PY
done
