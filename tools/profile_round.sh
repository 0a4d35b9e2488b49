#!/bin/bash
# One gpurun call: the plain bench (the number), then the ncu launch list of the same command
# and full captures of the sweep kernels. Outputs land in gpurun_out/ (copied to profiles/ here).
set -u
R=${1:-r1b}
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err || exit 1
$B > gpurun_out/plain_${R}.log 2>&1 || exit 1
# 3 warm-up steps x 47 launches are skipped; one timed step follows
ncu --metrics gpu__time_duration.sum --clock-control none -s 141 -c 47 --csv \
    --log-file gpurun_out/launches_${R}.csv $B > gpurun_out/ncu_list_${R}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_row2 -s 13 -c 2 \
    -o gpurun_out/prof_${R}_row $B > gpurun_out/ncu_row_${R}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_col -s 13 -c 1 \
    -o gpurun_out/prof_${R}_col $B > gpurun_out/ncu_col_${R}.log 2>&1
ls -la gpurun_out/ | grep ${R}
