#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_linalg_gpu.py tests/test_vlidort_gpu.py -m gpu -q -x > gpurun_out/r2i_tests1.log 2>&1; rc=$?; echo "quick tests rc=$rc"; tail -3 gpurun_out/r2i_tests1.log | cut -c1-250
if [ $rc -ne 0 ]; then grep -n "^E  " gpurun_out/r2i_tests1.log | head -20 | cut -c1-250; exit 0; fi
timeout -s KILL 900 python -m pytest tests/test_parity_baseline_gpu.py tests/test_pipeline_gpu.py -m gpu -q > gpurun_out/r2i_tests2.log 2>&1; echo "parity tests rc=$?"; tail -3 gpurun_out/r2i_tests2.log | cut -c1-200
timeout -s KILL 600 python bench.py --steps 3 --warmup 3 --no-aop --no-cpu --no-extras > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2i_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2i_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["bvp"]["frac"])
PY
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:k_vl_add -c 1 -o gpurun_out/r2i_full_add -f python bench.py --steps 1 --warmup 3 --no-aop --no-cpu --no-extras --nobs 1024 > gpurun_out/r2i_full_add.log 2>&1
echo done
