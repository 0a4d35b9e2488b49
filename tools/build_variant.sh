#!/bin/bash
# Builds lib/libpm_<tag>.so with extra -D flags on pm_sweep.cu (the other objects are reused);
# select it at run time with PM_B200_LIB=libpm_<tag>.so. usage: tools/build_variant.sh <tag> -DX=1 ...
set -e
T=$1; shift
D=ocean-perception_b200
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false \
  -Xcompiler -fPIC,-O3,-Wall "$@" -c $D/csrc/pm_sweep.cu -o $D/lib/pm_sweep_$T.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $D/lib/libpm_$T.so \
  $D/lib/pm_kernels.o $D/lib/pm_sweep_$T.o $D/lib/pm_cpu_semantics.o $D/lib/pm_seed.o $D/lib/pm_engine.o $D/lib/pm_yaml.o
rm -f $D/lib/pm_sweep_$T.o
echo built $D/lib/libpm_$T.so
