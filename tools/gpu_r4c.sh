#!/bin/bash
# the whole GPU suite (incl. the new cloud / NOBM forms) + smoke
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/r4c_tests.log 2>&1; echo "tests rc=$?"; tail -9 gpurun_out/r4c_tests.log | cut -c1-250
grep -n "^E  " gpurun_out/r4c_tests.log | head -10 | cut -c1-250
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4c_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r4c_smoke.log
