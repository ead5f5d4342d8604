#!/bin/bash
# new entry points: cloud variants, geometry calculators, C harness, plus the RT regression tests
mkdir -p gpurun_out
timeout -s KILL 500 python -m pytest tests/test_cloud_gpu.py tests/test_angles.py tests/test_cabi.py tests/test_vlidort_gpu.py tests/test_mie_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x > gpurun_out/r2o_tests.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -4 gpurun_out/r2o_tests.log | cut -c1-300
grep -n "^E  " gpurun_out/r2o_tests.log | head -20 | cut -c1-300
