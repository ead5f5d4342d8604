#!/bin/bash
# K3 phase timers on the GPU box (library built beforehand by tools/build_timing_lib.sh) -> gpurun_out/bvp_timing.json
mkdir -p gpurun_out
AB200_LIB=$PWD/aeroapps_b200/libaeroapps_b200_timing.so python tools/bvp_timing.py $1 > gpurun_out/bvp_timing.json 2> gpurun_out/bvp_timing.err; echo "timing rc=$?"
cat gpurun_out/bvp_timing.json; tail -3 gpurun_out/bvp_timing.err
