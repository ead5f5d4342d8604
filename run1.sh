set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
which gfortran f2c flang || true
./tools/fp64_peak | tee gpurun_out/fp64_peak.json
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 20 --warmup 5 2>&1 | tee gpurun_out/bench_aop.json | tail -2
python bench.py --impl reference --steps 3 --warmup 1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01_aop.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_aop_scalar -s 3 -c 2 -o gpurun_out/prof_aop_scalar python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/b2.log 2>&1
python - <<'PY' > gpurun_out/vec.log 2>&1
import time, numpy as np, torch, ctypes as C
from aeroapps_b200 import MieObs_, mietables as mt
tables = mt.synthetic_tables(nmom=300)
mie = MieObs_.ChemMie([550e-9], 300, tables)
st = mt.synthetic_state(nobs=512, km=72)
t0=time.time(); out = mie.getaopvector([550.0], st["vnames"], 0, st["qm"], st["rh"], 300, 6); t1=time.time()
t0=time.time(); out = mie.getaopvector([550.0], st["vnames"], 0, st["qm"], st["rh"], 300, 6); t1=time.time()
print("vector e2e 512 cols s:", t1-t0, "voxels/s", 512*72/(t1-t0))
PY
cat gpurun_out/vec.log
