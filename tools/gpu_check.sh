#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, launch list.  Run through gpurun from the repo root:
#   gpurun --timeout 1500 -- tools/gpu_check.sh
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests_gpu.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("value", d["value"], d["unit"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "cpu", d.get("cpu_baseline", {}).get("value"))
PY
