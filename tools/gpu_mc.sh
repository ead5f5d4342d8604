#!/bin/bash
mkdir -p gpurun_out
f=gpurun_out/mc_pytest.log
timeout -s KILL 600 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_vlidort_gpu.py -q -x -k "scalar_lambert and 8-False" > $f 2>&1; echo "memcheck pytest rc=$?"
grep -B2 -A16 "Invalid\|Misaligned\|misaligned" $f | head -70 | cut -c1-220; tail -n 4 $f | cut -c1-200
python tools/memcheck_case2.py 3 12 3 2>&1 | tail -3
AB200_VL_BVP_LEGACY=1 python tools/memcheck_case2.py 3 12 3 2>&1 | tail -3
