#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_linalg_gpu.py -m gpu -q -x > gpurun_out/r2h_linalg.log 2>&1; rc=$?; echo "linalg rc=$rc"; tail -4 gpurun_out/r2h_linalg.log | cut -c1-250
if [ $rc -ne 0 ]; then grep -n "^E  " gpurun_out/r2h_linalg.log | head -20 | cut -c1-250; exit 0; fi
timeout -s KILL 1500 python -m pytest tests -m gpu -q --deselect tests/test_linalg_gpu.py > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r2h_tests.log | cut -c1-200
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 --no-aop --no-cpu > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2h_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2h_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["kernel_ms_share"])
PY
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 1 --warmup 3 --no-aop --no-cpu --nobs 1024 > gpurun_out/r2h_launch.log 2>&1
python tools/launch_summary.py gpurun_out/r2h_launches.csv > gpurun_out/r2h_launch_summary.txt; head -6 gpurun_out/r2h_launch_summary.txt
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_vl_add -c 1 -o gpurun_out/r2h_full_add -f python bench.py --steps 1 --warmup 3 --no-aop --no-cpu --nobs 1024 > gpurun_out/r2h_full_add.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_vl_layer_ops -c 1 -o gpurun_out/r2h_full_ops -f python bench.py --steps 1 --warmup 3 --no-aop --no-cpu --nobs 1024 > gpurun_out/r2h_full_ops.log 2>&1
echo done
