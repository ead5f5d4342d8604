#!/bin/bash
# one GPU pass of the development loop: Stage A parity + bench (direct vs tiled kernel), Stage B parity + bench
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/r_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r_tests.log
for v in 1 0; do AB200_AOP_DIRECT=$v timeout -s KILL 300 python bench.py --workload aop_scalar --steps 20 --warmup 5 --no-cpu > gpurun_out/q_aop_$v.json 2> gpurun_out/q_aop_$v.err; tail -2 gpurun_out/q_aop_$v.err; python -c "
import json; d=json.load(open('gpurun_out/q_aop_$v.json')); print('aop direct=$v', d['value'], d['ms_per_step'], d['roofline']['frac'])"; done
timeout -s KILL 300 python bench.py --steps 3 --warmup 3 --no-aop --no-cpu > gpurun_out/q_bench_r.json 2> gpurun_out/q_bench_r.err; tail -2 gpurun_out/q_bench_r.err
python -c "
import json; d=json.load(open('gpurun_out/q_bench_r.json')); print('rt', d['value'], d['e2e']['value'], d['pipeline']['value'], d['roofline']['kernel_ms_share'])"
