#!/bin/bash
# Builds a variant of libnmcfs.so for A/B measurements: ab_build.sh NAME [wost_fast.cu path] [extra nvcc flags...]
# -> neural-monte-carlo-fluid-simulation_b200/build/variants/libnmcfs_NAME.so (git-ignored, travels with gpurun)
set -e
cd "$(dirname "$0")/../../neural-monte-carlo-fluid-simulation_b200"
NAME=$1; SRC=${2:-csrc/wost_fast.cu}; shift; shift || true
mkdir -p build/variants
cp "$SRC" csrc/_variant_wost_fast.cu
nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off -use_fast_math ${MINB:+-DNMC_MINB=$MINB} -DNMC_TRAV_INLINE "$@" -c csrc/_variant_wost_fast.cu -o build/variants/wost_fast_$NAME.o
rm csrc/_variant_wost_fast.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libnmcfs_$NAME.so build/wost_det.o build/variants/wost_fast_$NAME.o build/capi.o build/siren.o build/siren_tc.o build/fields.o build/peaks.o build/scene_build.o build/bessel_table.o
echo built build/variants/libnmcfs_$NAME.so
