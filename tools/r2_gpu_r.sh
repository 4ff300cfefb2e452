#!/bin/bash
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 200 python bench.py --steps 20 --warmup 3 --no-config4 > gpurun_out/r2r_$tag.json 2> gpurun_out/r2r_$tag.err
  rc=$?
  python - $tag $rc <<'PY'
import json,sys
try:
    l=[x for x in open('gpurun_out/r2r_%s.json' % sys.argv[1]) if x.startswith('{')][-1]
    d=json.loads(l); r=d['roofline']
    print(sys.argv[1], 'rc', sys.argv[2], 'config2 ms', round(d['ms_per_step'],4), 'scan ms', round(r['kernel_ms'],4), 'frac', round(r['frac'],3))
except Exception as e:
    print(sys.argv[1], 'rc', sys.argv[2], 'failed', e)
PY
}
run nc2_pair1 EDSB_FUSED_NC=2 EDSB_FUSED_PAIR=1
run nc2_pair0 EDSB_FUSED_NC=2 EDSB_FUSED_PAIR=0
run nc2_pair1_dw6 EDSB_FUSED_NC=2 EDSB_FUSED_PAIR=1 EDSB_FUSED_DW=6
run nc2_pair1_pw2 EDSB_FUSED_NC=2 EDSB_FUSED_PAIR=1 EDSB_FUSED_PW=2
run nc4_pair1 EDSB_FUSED_NC=4 EDSB_FUSED_PAIR=1
run nc1_pair1_pw2 EDSB_FUSED_PAIR=1 EDSB_FUSED_PW=2
