#!/bin/bash
# tools/build_variant.sh <name> <unit> [nvcc defines...]: library variant with one translation unit recompiled
# (development aid for A/B runs: FG_B200_LIB=figaroh-plus_b200/lib/variants/lib_<name>.so)
set -e
name=$1; unit=$2; shift 2
cd "$(dirname "$0")/.."
L=figaroh-plus_b200/lib; mkdir -p $L/variants $L/obj
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c -o $L/obj/${unit}_$name.o figaroh-plus_b200/csrc/$unit.cu
objs=""
for u in fg_api fg_gq fg_wide fg_matz fg_coop fg_rq; do
  if [ $u = $unit ]; then objs="$objs $L/obj/${unit}_$name.o"; else objs="$objs $L/obj/$u.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o $L/variants/lib_$name.so $objs
echo built $L/variants/lib_$name.so
