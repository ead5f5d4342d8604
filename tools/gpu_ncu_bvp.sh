#!/bin/bash
# one full ncu capture of the first k_vl_bvp launch (592 pixels = 4 per CTA); report lands in gpurun_out/
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:${1:-k_vl_bvp} --launch-count 1 -o gpurun_out/${2:-ncu_bvp} -f python bench.py --steps 1 --warmup 0 --no-cpu --no-aop --nobs 592 > gpurun_out/ncu_capture.log 2>&1; echo "ncu rc=$?"
