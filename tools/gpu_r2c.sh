#!/bin/bash
# round 2: RT parity tests, bench (default K3), launch list
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_vlidort_gpu.py tests/test_parity_baseline_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2c_tests.log
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 --no-aop --no-cpu > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2c_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2c_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["kernel_ms_share"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 1 --warmup 3 --no-aop --no-cpu --nobs 1024 > gpurun_out/r2c_launch.log 2>&1
python tools/launch_summary.py gpurun_out/r2c_launches.csv | head -6
