#!/bin/bash
# round 2, first GPU pass: parity (incl. the BASELINE-shape tests), smoke, bench with the new and the legacy K3
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2a_tests.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2a_smoke.log
for leg in 0 1; do
AB200_VL_BVP_LEGACY=$leg timeout -s KILL 600 python bench.py --steps 3 --warmup 3 --no-aop --no-cpu > gpurun_out/r2a_bench_leg$leg.json 2> gpurun_out/r2a_bench_leg$leg.err; echo "bench legacy=$leg rc=$?"; tail -2 gpurun_out/r2a_bench_leg$leg.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2a_bench_leg$leg.json"))
print("legacy=$leg value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["kernel_ms_share"])
PY
done
timeout -s KILL 600 python bench.py --workload rt_scalar --steps 3 --warmup 3 --no-cpu > gpurun_out/r2a_bench_scalar.json 2> gpurun_out/r2a_bench_scalar.err; echo "scalar rc=$?"
AB200_VL_BVP_LEGACY=1 timeout -s KILL 600 python bench.py --workload rt_scalar --steps 3 --warmup 3 --no-cpu > gpurun_out/r2a_bench_scalar_leg.json 2> gpurun_out/r2a_bench_scalar_leg.err
python - <<PY
import json
for f in ("r2a_bench_scalar", "r2a_bench_scalar_leg"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, "value", d["value"], d["roofline"]["kernel_ms_share"])
PY
