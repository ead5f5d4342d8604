#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_bpdf_gpu.py -m gpu -q -s > gpurun_out/r3b_tests.log 2>&1; echo "rc=$?"; grep -n "parity \|passed\|failed\|^E  " gpurun_out/r3b_tests.log | cut -c1-700 | head -20
