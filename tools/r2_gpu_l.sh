#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_group.py tests/test_cli.py -x -q -m gpu 2>&1 | tail -4
cat > /tmp/gl.py <<'PY'
import sys, time, json
sys.path.insert(0,'tests'); sys.path.insert(0,'tools'); sys.path.insert(0,'.')
import edsparser_b200 as E, bench_leds
lib=E.load()
e,s=bench_leds.genrandomeds_like(100_000_000)
c=lib.context(0)
one=c.leds_merge_host(e,s,10)
t0=time.perf_counter(); one=c.leds_merge_host(e,s,10); t1=time.perf_counter()-t0
c.close()
for n in (2,4):
    g=lib.group(list(range(n)))
    got=g.leds_merge_host(e,s,10)
    t0=time.perf_counter(); got=g.leds_merge_host(e,s,10); t=time.perf_counter()-t0
    print(json.dumps({"config3_100Mbp_group":n,"shards_used":got[3],"equal_to_single_device":got[0]==one[0] and got[1]==one[1],"host_to_host_s":round(t,3),"single_device_s":round(t1,3)}),flush=True)
    g.close()
PY
timeout 600 python /tmp/gl.py > gpurun_out/r2l_group_leds.jsonl 2>&1; cat gpurun_out/r2l_group_leds.jsonl | tail -4
