#!/bin/bash
# launch list of the bench command at HEAD (after the plain run has exited 0), then one --set full capture of the scan
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-config4"
timeout 300 $CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_h_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02_h_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_scan_fused -s 4 -c 1 -o gpurun_out/r02_h_scan_prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
ncu -i gpurun_out/r02_h_scan_prof.ncu-rep --page raw --csv > gpurun_out/r02_h_k_scan_fused_ncu_full.csv 2>/dev/null
wc -c gpurun_out/r02_h_k_scan_fused_ncu_full.csv
