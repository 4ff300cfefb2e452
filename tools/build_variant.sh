#!/bin/bash
# tools/build_variant.sh <name> <extra nvcc flags...>: the product library with other compile-time constants, for A/B runs
# (EDSB_LIBRARY=build/variants/lib<name>.so python tools/bench_leds.py ...). Not part of the product build.
set -e
name=$1; shift
mkdir -p build/variants/$name
for f in msa leds vcf capi shard; do
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" -c edsparser_b200/csrc/$f.cu -o build/variants/$name/$f.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/variants/lib$name.so build/variants/$name/*.o -ldl
rm -rf build/variants/$name
ls -la build/variants/lib$name.so
