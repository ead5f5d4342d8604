#!/bin/bash
# multi-geometry moment sums: parity of the 60-view calls + the RT regression tests, then the polar16 and headline benches
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_parity_baseline_gpu.py tests/test_vlidort_gpu.py -m gpu -q -x -k "multigeom or multi_chunk or polar or lambert" > gpurun_out/r3a_tests.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -3 gpurun_out/r3a_tests.log | cut -c1-250
if [ $rc -ne 0 ]; then grep -n "^E  " gpurun_out/r3a_tests.log | head -20 | cut -c1-250; exit 0; fi
timeout -s KILL 200 python bench.py --workload polar16 --steps 2 --warmup 3 > gpurun_out/r3a_polar.json 2> gpurun_out/r3a_polar.err; echo "polar rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r3a_polar.json"))
print("polar16 value", d["value"], "ms", d["ms_per_step"], d["roofline"]["kernel_ms"], "parity", d.get("parity", {}).get("pass"), d.get("parity", {}).get("max_rel"))
PY
