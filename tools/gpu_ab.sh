#!/bin/bash
# A/B of the two K3 forward eliminations (AB200_VL_BVP_V3=1 warp-specialised, =0 blocked look-ahead) after smoke + parity tests
mkdir -p gpurun_out
timeout -s KILL 240 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/q_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/q_smoke.log
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/q_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/q_tests.log
for v in 1 0; do
AB200_VL_BVP_V3=$v timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-aop --no-cpu > gpurun_out/q_bench_v3_$v.json 2> gpurun_out/q_bench_v3_$v.err; echo "bench v3=$v rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/q_bench_v3_$v.json')); print(d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['kernel_ms_share'], d['clocks'])"
done
nproc; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
