#!/bin/bash
# scratch GPU script: parity tests, bench, launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r5_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r5_tests.log
tail -5 gpurun_out/r5_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r5_bench.json 2> gpurun_out/r5_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r5_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r5_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-aop > gpurun_out/r5_ncu_bench.log 2>&1; echo "ncu rc=$?"
