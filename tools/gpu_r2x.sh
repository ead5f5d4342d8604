#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests/test_vlidort_gpu.py tests/test_parity_baseline_gpu.py tests/test_cloud_gpu.py tests/test_cabi.py tests/test_pipeline_gpu.py -m gpu -q -x > gpurun_out/r2x_tests.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -3 gpurun_out/r2x_tests.log | cut -c1-250
if [ $rc -ne 0 ]; then grep -n "^E  " gpurun_out/r2x_tests.log | head -20 | cut -c1-250; exit 0; fi
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-aop --no-cpu --no-extras > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2x_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2x_bench.json"))
print("value", d["value"], "e2e", d["e2e"], "ms", d["ms_per_step"])
PY
