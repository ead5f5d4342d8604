#!/bin/bash
# launch list of one bench step + full captures of the new K2b / K3 kernels (m = 0 launch)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 1 --warmup 3 --no-aop --no-cpu --nobs 1024 > gpurun_out/r2b_launch.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_vl_add -c 1 -o gpurun_out/r2b_full_add -f python bench.py --steps 1 --warmup 3 --no-aop --no-cpu --nobs 1024 > gpurun_out/r2b_full_add.log 2>&1
echo "full add rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_vl_layer_ops -c 1 -o gpurun_out/r2b_full_ops -f python bench.py --steps 1 --warmup 3 --no-aop --no-cpu --nobs 1024 > gpurun_out/r2b_full_ops.log 2>&1
echo "full ops rc=$?"
