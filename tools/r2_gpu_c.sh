#!/bin/bash
mkdir -p gpurun_out
EDSB_FUSED_MODE=0 python tools/profile_msa.py 100 10000000 3 > gpurun_out/r2c_plain.log 2>&1 &&
EDSB_FUSED_MODE=0 ncu --set full --clock-control none --import-source on -k regex:k_scan_fused -s 1 -c 1 -o gpurun_out/r2c_fused_bulk python tools/profile_msa.py 100 10000000 3 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
