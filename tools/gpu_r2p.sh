#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_cloud_gpu.py -m gpu -q -s > gpurun_out/r2p_tests.log 2>&1; echo "rc=$?"; grep -n "parity \|passed\|failed" gpurun_out/r2p_tests.log | cut -c1-900
