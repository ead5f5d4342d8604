#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_vlidort_gpu.py tests/test_parity_baseline_gpu.py tests/test_bpdf_gpu.py tests/test_cloud_gpu.py -m gpu -q -x > gpurun_out/r3c_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r3c_tests.log | cut -c1-200
grep -n "^E  " gpurun_out/r3c_tests.log | head -5 | cut -c1-300
