#!/bin/bash
# scripts/build_variant.sh <name> <extra nvcc flags...>: an experiment build of libp2e.so into scratch/lib_<name>.so
# (bench.py / tests pick it up with P2E_LIB=scratch/lib_<name>.so); the product build is plonky2_ecdsa_b200/build.py
set -e
NAME=$1; shift
cd "$(dirname "$0")/.."
mkdir -p scratch/obj_$NAME
C=plonky2_ecdsa_b200/csrc
for f in p2e_kernels.cu p2e_chain.cu p2e_ux.cu p2e_ntt.cu p2e_genstream.cu p2e_tape.cpp; do
  o=scratch/obj_$NAME/${f%.*}.o
  if [ "$f" = "p2e_chain.cu" ] || [ ! -f $o ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -x cu -c $C/$f -o $o &
  fi
done
wait
nvcc -shared -cudart shared -o scratch/lib_$NAME.so scratch/obj_$NAME/*.o -gencode arch=compute_100a,code=sm_100a -Xlinker -rpath=/usr/local/cuda/lib64
echo built scratch/lib_$NAME.so
