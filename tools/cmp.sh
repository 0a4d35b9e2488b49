B="python bench.py --steps 1 --warmup 3 --pairs-per-gpu 16 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain.log 2>&1 || exit 1
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section WarpStateStats --section LaunchStats --section Occupancy --section SchedulerStats"
PM_SWEEP_V1=1 ncu $SEC --clock-control none -k regex:k_sweep_col -s 4 -c 1 -o gpurun_out/cmp_v1 $B > gpurun_out/cmp_v1.log 2>&1
PM_COL_VAR=0 ncu $SEC --clock-control none -k regex:k_sweep_col2 -s 4 -c 1 -o gpurun_out/cmp_v2a $B > gpurun_out/cmp_v2a.log 2>&1
PM_COL_VAR=4 ncu $SEC --clock-control none -k regex:k_sweep_col2 -s 4 -c 1 -o gpurun_out/cmp_v2b $B > gpurun_out/cmp_v2b.log 2>&1
ls -la gpurun_out/cmp_*
