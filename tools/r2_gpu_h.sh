#!/bin/bash
mkdir -p gpurun_out
timeout 100 python tools/profile_msa.py 1000 3000000 3 > gpurun_out/r2h_plain.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_emit_var|k_group3|k_emit3|k_group$" -s 4 -c 4 -o gpurun_out/r2h_sym python tools/profile_msa.py 1000 3000000 3 > gpurun_out/r2h_ncu.log 2>&1
tail -3 gpurun_out/r2h_ncu.log
