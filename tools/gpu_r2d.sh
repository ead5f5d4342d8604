#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests/test_vlidort_gpu.py -m gpu -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2d_tests.log | cut -c1-200
cat > /tmp/race_case.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np
from aeroapps_b200 import VLIDORT_POLAR_ as VP
from oracle import vlidort_glue as vg
sc = vg.synthetic_scene(3, nmom=16, seed=48)
I, R, rc = VP.scalar_lambert(sc["channels"], 8, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["g"], sc["pmom"], sc["pe"], sc["he"], sc["te"], sc["albedo"], sc["sza"], sc["raa"], sc["vza"], 1e15, 0)
ref = vg.vlidort_run(1, 8, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["he"], sc["sza"], sc["raa"], sc["vza"], albedo=sc["albedo"], scalar_glue=True)
print("rel", np.max(np.abs(I - ref["radiance"]) / ref["radiance"]))
PY
timeout -s KILL 600 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/race_case.py > gpurun_out/r2d_race.log 2>&1; echo "racecheck rc=$?"; grep -v "^=========     at\|^=========     in" gpurun_out/r2d_race.log | tail -40 | cut -c1-220
