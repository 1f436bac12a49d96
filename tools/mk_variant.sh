#!/bin/bash
# Build a library variant for tools/ab.sh:  tools/mk_variant.sh NAME [-DMACRO=VALUE ...]
#   -> sdr-j-dab_b200/variants/lib_NAME.so (git-ignored; travels to the GPU box with the snapshot)
# The macros are the kernels' A/B switches (VS_NACC, VS_FDEC, VS_PL2, VS_CW, R8_TMA, R8_TWPOW, R8_MINB, TB_THREADS, TB_NBUF ...).
set -e
cd "$(dirname "$0")/../sdr-j-dab_b200"
name=$1; shift
mkdir -p variants/obj_$name
for f in csrc/*.cu csrc/*.cpp; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" -c -o variants/obj_$name/$(basename $f).o $f 2>/dev/null &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/lib_$name.so variants/obj_$name/*.o
rm -rf variants/obj_$name
ls -la variants/lib_$name.so
