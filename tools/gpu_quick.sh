#!/bin/bash
# quick GPU pass: smoke, parity tests, short bench (each under its own timeout so a hung kernel cannot eat the box)
mkdir -p gpurun_out
timeout -s KILL 240 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/q_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/q_smoke.log
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/q_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/q_tests.log
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-aop --no-cpu > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/q_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['kernel_ms_share'])"
