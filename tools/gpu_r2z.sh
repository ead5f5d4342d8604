#!/bin/bash
# the driver's scaling launch at N = 4: reference arm, then the default bench line under torchrun
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
( time timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > gpurun_out/r2z_ref_n4.json 2> gpurun_out/r2z_ref_n4.err ) 2>&1 | grep real; echo "ref rc=$?"; tail -c 300 gpurun_out/r2z_ref_n4.json
( time timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2z_bench_n4.json 2> gpurun_out/r2z_bench_n4.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/r2z_bench_n4.err | cut -c1-300
python - <<PY
import json
d = json.load(open("gpurun_out/r2z_bench_n4.json"))
print("n4 value", d["value"], "e2e", d["e2e"]["value"], d["config"].get("parallelism"))
x = d.get("in_library_multi_gpu"); print("in-library", x and (x["value"], x["devices"], x["ms_per_call"]))
for k, x in d.get("extras", {}).items(): print(k, x["value"])
PY
