#!/bin/bash
# 2 GPUs: full GPU test-suite (new: C harness, float32 table upload, multi-chunk, pipeline with a non-float32 wavelength) + 2-rank bench
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests/test_multidevice_gpu.py tests/test_vlidort_gpu.py tests/test_cloud_gpu.py -m gpu -q -x > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"; tail -14 gpurun_out/r2w_tests.log | cut -c1-250
grep -n "^E  " gpurun_out/r2w_tests.log | head -10 | cut -c1-250
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2w_bench_n2.json 2> gpurun_out/r2w_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r2w_bench_n2.err | cut -c1-300
python - <<PY
import json
d = json.load(open("gpurun_out/r2w_bench_n2.json"))
print("n2 value", d["value"], "e2e", d["e2e"]["value"], d["config"].get("parallelism"))
print("in-library", d.get("in_library_multi_gpu"))
for k, x in d.get("extras", {}).items(): print(k, x["value"])
PY
