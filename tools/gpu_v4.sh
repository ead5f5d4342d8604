#!/bin/bash
# K3 register-resident forward elimination: parity tests with it forced on for every n, then the bench with / without it
mkdir -p gpurun_out
AB200_VL_BVP_V4=2 timeout -s KILL 900 python -m pytest tests/test_vlidort_gpu.py -m gpu -x -q > gpurun_out/v4_tests.log 2>&1; echo "tests(v4 forced) rc=$?"; tail -15 gpurun_out/v4_tests.log
for v in 1 0; do
AB200_VL_BVP_V4=$v timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-aop --no-cpu > gpurun_out/q_bench_v4_$v.json 2> gpurun_out/q_bench_v4_$v.err; echo "bench v4=$v rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/q_bench_v4_$v.json')); print(d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['kernel_ms_share'])"
done
