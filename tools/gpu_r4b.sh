#!/bin/bash
# final ncu evidence of the round (kernels as committed): launch list of the bench command + one full capture of each Stage-B kernel
# (Fourier term m = 1 launch, 1024 pixels) + the Stage A scalar kernel
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r4b_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-aop --no-extras > gpurun_out/r4b_launches.log 2>&1; echo "launch list rc=$?"
for k in k_vl_eigen k_vl_layer_ops k_vl_add; do
  timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 1 -c 1 -o gpurun_out/r4b_full_$k -f python bench.py --steps 1 --warmup 0 --no-aop --no-cpu --no-extras --nobs 1024 > gpurun_out/r4b_full_$k.log 2>&1; echo "$k rc=$?"
done
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:k_aop_scalar --launch-skip 1 -c 1 -o gpurun_out/r4b_full_k_aop_scalar -f python bench.py --workload aop_scalar --steps 1 --warmup 1 --no-cpu > gpurun_out/r4b_full_k_aop_scalar.log 2>&1; echo "aop rc=$?"
ls -la gpurun_out/r4b_*
