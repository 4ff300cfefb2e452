#!/bin/bash
# fused scan with tiles fetched in pairs: parity first (short timeouts: a pipeline bug hangs), then the A/B bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_msa_gpu.py -q -m gpu -x -k "fused or synth or random" > gpurun_out/r2q_pytest.log 2>&1
rc=$?
echo "pytest rc=$rc" >> gpurun_out/r2q_pytest.log
tail -4 gpurun_out/r2q_pytest.log
if [ $rc -ne 0 ]; then exit 1; fi
for pair in 1 0; do
  EDSB_FUSED_PAIR=$pair timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2q_bench_pair$pair.json 2> gpurun_out/r2q_bench_pair$pair.err
  echo "bench pair=$pair rc=$?"
  python - $pair <<'PY'
import json,sys
l=[x for x in open('gpurun_out/r2q_bench_pair%s.json' % sys.argv[1]) if x.startswith('{')][-1]
d=json.loads(l)
r=d['roofline']
print('config2 ms', d['ms_per_step'], 'scan ms', r['kernel_ms'], 'frac', round(r['frac'],3), 'e2e', d['e2e']['value'])
c=d.get('config4')
if c: print('config4 ms', c['ms_per_step'], 'scan', c['roofline']['kernel_ms'], 'frac', round(c['roofline']['frac'],3))
PY
done
timeout 600 python -m pytest tests/test_msa_gpu.py tests/test_full_size_gpu.py -q -m gpu -x -k "msa" > gpurun_out/r2q_pytest2.log 2>&1
echo "pytest2 rc=$?"; tail -3 gpurun_out/r2q_pytest2.log
