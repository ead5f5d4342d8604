#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_vl_bvp --launch-count 1 -o gpurun_out/r8_bvp -f python bench.py --steps 1 --warmup 0 --no-cpu --no-aop --nobs 592 > gpurun_out/r8_ncu_bvp.log 2>&1; echo "ncu bvp rc=$?"
