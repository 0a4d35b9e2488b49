#!/bin/bash
# Round-2 profile pass (one gpurun call): the plain bench, the ncu launch list of one device pass,
# DRAM bytes of every sweep launch, full captures of the sweep kernels at level 0.
# Outputs land in gpurun_out/ (summarised into profiles/ with tools/ncu_summary.py).
# usage: tools/profile_r2.sh <tag> [env assignments...]
set -u
R=$1; shift
# one device pass of 64 pairs per step: the launch structure of the 512-pair bench, 1/8 of its length
B="python bench.py --pairs-per-gpu 64 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
env "$@" $B > gpurun_out/plain_${R}.log 2>&1 || { tail -5 gpurun_out/plain_${R}.log; exit 1; }
# launches per step: 1 (fp32 probe is outside) ... count them from the list itself
env "$@" ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_${R}.csv $B > gpurun_out/ncu_list_${R}.log 2>&1
# DRAM traffic of the sweep launches (all steps; the summary keeps the last 24 = the timed step)
env "$@" ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:k_sweep --csv --log-file gpurun_out/sweep_dram_${R}.csv $B > gpurun_out/ncu_dram_${R}.log 2>&1
# level-0 launches of the timed step: a step runs 6 level-1 row sweeps, then 6 level-0 ones
env "$@" ncu --set full --clock-control none --import-source on -k regex:k_sweep_row -s 42 -c 2 \
    -o gpurun_out/prof_${R}_row $B > gpurun_out/ncu_row_${R}.log 2>&1
env "$@" ncu --set full --clock-control none --import-source on -k regex:k_sweep_col -s 42 -c 2 \
    -o gpurun_out/prof_${R}_col $B > gpurun_out/ncu_col_${R}.log 2>&1
# the streaming kernels of the timed step: DRAM bytes and time per launch (HBM fraction of each)
env "$@" ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'k_preprocess|k_mask_background|k_finalize|k_upsample2|k_init_random|k_downscale2' --csv \
    --log-file gpurun_out/stream_dram_${R}.csv $B > gpurun_out/ncu_stream_${R}.log 2>&1
ls -la gpurun_out/ | grep ${R}
