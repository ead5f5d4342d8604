#!/bin/bash
# checkpoint: the whole GPU test-suite, smoke(), then the driver's round-end sequence (reference arm, default bench line)
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"; tail -9 gpurun_out/r2q_tests.log | cut -c1-250
grep -n "^E  " gpurun_out/r2q_tests.log | head -10 | cut -c1-250
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2q_smoke.log
bash tools/gpu_bench.sh
