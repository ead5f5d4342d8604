#!/bin/bash
# pipelined slabs on one device: regression + parity at the BASELINE shapes (256 observations -> 3 slabs), then the headline bench
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_vlidort_gpu.py tests/test_parity_baseline_gpu.py tests/test_cloud_gpu.py tests/test_cabi.py -m gpu -q -x > gpurun_out/r2r_tests.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -3 gpurun_out/r2r_tests.log | cut -c1-250
if [ $rc -ne 0 ]; then grep -n "^E  " gpurun_out/r2r_tests.log | head -20 | cut -c1-250; exit 0; fi
for v in 1 3; do
AB200_PIPELINE_SLABS=$v timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-aop --no-cpu --no-extras > gpurun_out/r2r_bench_$v.json 2> gpurun_out/r2r_bench_$v.err; echo "bench slabs=$v rc=$?"; tail -2 gpurun_out/r2r_bench_$v.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2r_bench_$v.json"))
print("slabs $v value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("pageable_host_value"), "ms", d["ms_per_step"], d.get("pipeline", {}).get("value"))
PY
done
