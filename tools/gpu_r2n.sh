#!/bin/bash
# one full ncu capture of the K2 eigen kernel (Fourier term m = 1 launch, 1024 pixels)
mkdir -p gpurun_out
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_vl_eigen --launch-skip 1 -c 1 -o gpurun_out/r2n_full_eigen -f python bench.py --steps 1 --warmup 0 --no-aop --no-cpu --no-extras --nobs 1024 > gpurun_out/r2n_full_eigen.log 2>&1; echo "rc=$?"
