#!/bin/bash
# K2 convergence-flag variant: RT parity tests + the headline line without the CPU legs
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_vlidort_gpu.py tests/test_parity_baseline_gpu.py tests/test_max_sizes_gpu.py -m gpu -q -x > gpurun_out/r4f_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r4f_tests.log | cut -c1-200
timeout -s KILL 200 python bench.py --no-cpu --no-aop --no-extras > gpurun_out/r4f_bench.json 2> gpurun_out/r4f_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r4f_bench.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
