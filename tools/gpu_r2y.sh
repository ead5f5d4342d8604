#!/bin/bash
mkdir -p gpurun_out
for v in 1 2 3; do
AB200_PIPELINE_SLABS=$v timeout -s KILL 200 python bench.py --steps 2 --warmup 3 --no-aop --no-cpu --no-extras > gpurun_out/r2y_bench_$v.json 2> gpurun_out/r2y_bench_$v.err; echo "slabs=$v rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r2y_bench_$v.json"))
print("slabs $v value", round(d["value"]), "e2e pinned", round(d["e2e"]["value"]), "pageable", round(d["e2e"]["pageable_host_value"]))
PY
done
