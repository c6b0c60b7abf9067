#!/bin/bash
# build_variant.sh NAME "-DFLAG ..." : libpsc_b200_NAME.so with push.cu compiled with extra flags
# (A/B experiments on the GPU box: PSC_B200_LIB=libpsc_b200_NAME.so python bench.py ...)
set -e
cd "$(dirname "$0")/../psc_b200/csrc"
NAME=$1; EXTRA=$2
B=build/var_$NAME; mkdir -p $B
FL="-gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr -Xcudafe --diag_suppress=177"
nvcc $FL -fmad=false -DPUSH_VARIANT=exact $EXTRA -Xptxas -v -c push.cu -o $B/push_exact.o 2> $B/exact.log &
nvcc $FL -fmad=true -DPUSH_VARIANT=fast $EXTRA -c push.cu -o $B/push_fast.o 2> $B/fast.log &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libpsc_b200_$NAME.so build/capi.o build/particles.o $B/push_exact.o $B/push_fast.o build/sort.o build/bndp.o build/fields.o build/comm.o build/fused_sort.o build/collision.o build/checkpoint.o -ldl
grep -A2 "Function properties for _ZN8psc_b2005exact12k_push_tiledILi0ELi1ENS0_9GeoStaticILi0EEELb1ELb1ELi256ELi3" $B/exact.log | tail -2
