#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_max_sizes_gpu.py -m gpu -q -s --durations=3 > gpurun_out/r4e_tests.log 2>&1; echo "rc=$?"
grep -n "^parity\|passed\|failed\|^E  \|Error" gpurun_out/r4e_tests.log | cut -c1-500 | head -20
