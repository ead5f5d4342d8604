#!/bin/bash
mkdir -p gpurun_out
AB200_TRACE=1 AB200_PIPELINE_SLABS=3 timeout -s KILL 300 python bench.py --steps 2 --warmup 3 --no-aop --no-cpu --no-extras > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "rc=$?"
grep "ab200" gpurun_out/r2s_bench.err | tail -60
