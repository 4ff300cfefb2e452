#!/bin/bash
mkdir -p gpurun_out
free -g | head -2
timeout 1500 python tools/bench_vcf.py 1000000 > gpurun_out/r2k_vcf_config5.jsonl 2> gpurun_out/r2k_vcf_config5.err
echo "rc=$?"; tail -c 3000 gpurun_out/r2k_vcf_config5.jsonl; tail -5 gpurun_out/r2k_vcf_config5.err
