#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r6_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r6_tests.log
tail -4 gpurun_out/r6_tests.log
python bench.py --steps 3 --warmup 3 --no-aop --no-cpu > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r6_bench.json')); print(d['value'], d['e2e']['value'], d['roofline'])"
