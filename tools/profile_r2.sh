#!/bin/bash
# ncu captures of the sweep kernels at level 0 (one gpurun call). Outputs in gpurun_out/.
# usage: tools/profile_r2.sh <tag> [env assignments...]
set -u
R=$1; shift
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
env "$@" $B > gpurun_out/plain_${R}.log 2>&1 || { tail -5 gpurun_out/plain_${R}.log; exit 1; }
# level-0 launches: a step runs 6 level-1 row sweeps, then 6 level-0 ones; 3 warm-up steps + 1 init
# launch 42 = first level-0 row sweep of the timed step (noise-fused, +1), 43 = the plain -1 one
env "$@" ncu --set full --clock-control none --import-source on -k regex:k_sweep_row2 -s 42 -c 2 \
    -o gpurun_out/prof_${R}_row $B > gpurun_out/ncu_row_${R}.log 2>&1
env "$@" ncu --set full --clock-control none --import-source on -k regex:k_sweep_col -s 42 -c 2 \
    -o gpurun_out/prof_${R}_col $B > gpurun_out/ncu_col_${R}.log 2>&1
ls -la gpurun_out/ | grep ${R}
