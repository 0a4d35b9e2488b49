set -u
ncu --set full --clock-control none --import-source on -k regex:k_sweep_col -s 18 -c 3 -o gpurun_out/prof_r2c_c5_rowT python tools/band_bench.py --steps 1 --warmup 3 > gpurun_out/ncu_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_generic -s 18 -c 2 -o gpurun_out/prof_r2c_c5_gen python tools/band_bench.py --steps 1 --warmup 3 > gpurun_out/ncu_c5g.log 2>&1
ls -la gpurun_out/prof_r2c_c5*
