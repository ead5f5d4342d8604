#!/bin/bash
# quick A/B of a K2 change: RT tests (short timeout: a hung kernel must not eat the budget), then the headline bench without the CPU legs
mkdir -p gpurun_out
timeout -s KILL 150 python -m pytest tests/test_vlidort_gpu.py tests/test_linalg_gpu.py -m gpu -q -x > gpurun_out/r2m_tests.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -3 gpurun_out/r2m_tests.log | cut -c1-250
if [ $rc -ne 0 ]; then grep -n "^E  " gpurun_out/r2m_tests.log | head -20 | cut -c1-250; exit 0; fi
timeout -s KILL 200 python bench.py --steps 3 --warmup 3 --no-aop --no-cpu --no-extras > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2m_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2m_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["kernel_ms"])
PY
timeout -s KILL 200 python bench.py --workload polar16 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2m_polar.json 2> gpurun_out/r2m_polar.err; echo "polar rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r2m_polar.json"))
print("polar16 value", d["value"], "ms", d["ms_per_step"], d["roofline"]["kernel_ms"])
PY
