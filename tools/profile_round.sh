#!/bin/bash
# One gpurun call: the plain benches (the numbers), then the ncu launch list of the same command,
# DRAM bytes of every sweep launch of one step, and full captures of the sweep kernels at level 0.
# Outputs land in gpurun_out/ (summarised into profiles/ with tools/ncu_summary.py).
set -u
R=${1:-r1c}
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err || exit 1
python bench.py --init sparse --levels 1 --steps 5 --warmup 3 --cpu-sample-pairs 2 \
    > gpurun_out/bench_${R}_sparse.json 2> gpurun_out/bench_${R}_sparse.err || exit 1
$B > gpurun_out/plain_${R}.log 2>&1 || exit 1
# 3 warm-up steps x 47 launches are skipped; one timed step follows
ncu --metrics gpu__time_duration.sum --clock-control none -s 141 -c 47 --csv \
    --log-file gpurun_out/launches_${R}.csv $B > gpurun_out/ncu_list_${R}.log 2>&1
# DRAM traffic of the 24 sweep launches of the timed step (3 x 24 warm-up launches skipped)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:k_sweep -s 72 -c 24 --csv --log-file gpurun_out/sweep_dram_${R}.csv $B \
    > gpurun_out/ncu_dram_${R}.log 2>&1
# level-0 launches: a step runs 6 level-1 row sweeps, then 6 level-0 ones (12 column sweeps likewise)
ncu --set full --clock-control none --import-source on -k regex:k_sweep_row2 -s 42 -c 2 \
    -o gpurun_out/prof_${R}_row $B > gpurun_out/ncu_row_${R}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_col -s 42 -c 1 \
    -o gpurun_out/prof_${R}_col $B > gpurun_out/ncu_col_${R}.log 2>&1
# the seeding kernels (reference-default params)
S="python bench.py --init sparse --levels 1 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_seed -s 15 -c 5 --csv \
    --log-file gpurun_out/launches_${R}_seed.csv $S > gpurun_out/ncu_seed_${R}.log 2>&1
ls -la gpurun_out/ | grep ${R}
