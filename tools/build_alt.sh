#!/bin/bash
# Development builds of one translation unit with extra -D flags, linked against the default objects into tools/alt/libtfhe_b200_<name>.so
# (selected at run time with TFHE_B200_LIB=...).   Usage: tools/build_alt.sh <name> <tu: br_kernels|ks_kernels|...> "<flags>"
set -e
cd "$(dirname "$0")/../experimental-tfhe_b200"
NAME=$1; TU=$2; FLAGS=$3
mkdir -p ../tools/alt build
nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
     -Xptxas -v $FLAGS -c csrc/${SRCTU:-$TU}.cu -o build/alt_${NAME}_$TU.o 2> build/alt_${NAME}_$TU.ptxas.log || (cat build/alt_${NAME}_$TU.ptxas.log; false)
OBJS=""
for o in capi br_kernels ks_kernels ks_tc_kernels misc_kernels hp_kernels exact_kernels keygen_kernels twiddles; do
  if [ "$o" == "$TU" ]; then OBJS="$OBJS build/alt_${NAME}_$TU.o"; else OBJS="$OBJS build/$o.o"; fi
done
nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -shared -o ../tools/alt/libtfhe_b200_$NAME.so $OBJS -lquadmath -cudart static
grep -A2 "$4" build/alt_${NAME}_$TU.ptxas.log | grep -E "registers|spill" | head -4 || true
