#!/bin/bash
# new cloud forms: their tests + the older cloud / BPDF tests they lean on
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_cloud_more_gpu.py tests/test_cloud_gpu.py tests/test_bpdf_gpu.py -m gpu -q --durations=5 -s > gpurun_out/r4a_tests.log 2>&1; echo "tests rc=$?"
grep -n "^parity\|passed\|failed\|^E  " gpurun_out/r4a_tests.log | cut -c1-400 | head -40
