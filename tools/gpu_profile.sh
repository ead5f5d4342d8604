#!/bin/bash
# round profile: launch list of the bench command + one full capture of each RT kernel (m = 0 launch, 2048 pixels)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_rt.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-aop > gpurun_out/launches_rt.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k_vl_bvp --launch-count 1 -o gpurun_out/full_bvp -f python bench.py --steps 1 --warmup 0 --no-cpu --no-aop > gpurun_out/full_bvp.log 2>&1; echo "bvp rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k_vl_eigen --launch-count 1 -o gpurun_out/full_eig -f python bench.py --steps 1 --warmup 0 --no-cpu --no-aop > gpurun_out/full_eig.log 2>&1; echo "eig rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_aop_scalar --launch-count 1 -o gpurun_out/full_aop -f python bench.py --workload aop_scalar --steps 1 --warmup 1 --no-cpu > gpurun_out/full_aop.log 2>&1; echo "aop rc=$?"
