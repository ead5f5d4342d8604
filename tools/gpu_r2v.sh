#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_cloud_gpu.py tests/test_linalg_gpu.py tests/test_vlidort_gpu.py -m gpu -q -x > gpurun_out/r2v_tests.log 2>&1; rc=$?; echo "tests rc=$rc"; tail -3 gpurun_out/r2v_tests.log | cut -c1-300
grep -n "^E  \|parity scalar" gpurun_out/r2v_tests.log | head -10 | cut -c1-400
SAN_N=12 timeout -s KILL 420 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 5 python tools/sanitize_case.py > gpurun_out/r2v_race_12.log 2>&1; echo "racecheck rc=$?"; grep "RACECHECK SUMMARY\|^rc " gpurun_out/r2v_race_12.log
grep "Race reported" gpurun_out/r2v_race_12.log | sed 's/+0x[0-9a-f]*//g' | sort | uniq -c | head -5 | cut -c1-250
