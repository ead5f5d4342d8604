#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_vl_bvp --launch-count 1 -o gpurun_out/r7_bvp -f python bench.py --steps 1 --warmup 0 --no-cpu --no-aop --nobs 592 > gpurun_out/r7_ncu_bvp.log 2>&1; echo "ncu bvp rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_vl_eigen --launch-count 1 -o gpurun_out/r7_eig -f python bench.py --steps 1 --warmup 0 --no-cpu --no-aop --nobs 592 > gpurun_out/r7_ncu_eig.log 2>&1; echo "ncu eig rc=$?"
ls -la gpurun_out/*.ncu-rep
