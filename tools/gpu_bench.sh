#!/bin/bash
# the driver's round-end sequence: reference arm, then the default bench line
mkdir -p gpurun_out
( time timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real; echo "ref rc=$?"; tail -2 gpurun_out/bench_ref.err | cut -c1-300
( time timeout -s KILL 1500 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/bench_full.err | cut -c1-300
python - <<PY
import json
r = json.load(open("gpurun_out/bench_ref.json")); print("reference", r["value"], r["cpu_baseline"]["sample"][:80])
d = json.load(open("gpurun_out/bench_full.json"))
print("value", d["value"], "e2e", d["e2e"], "ms", d["ms_per_step"]); print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["bvp"])
print("cpu", d.get("cpu_baseline")); print("parity", d.get("parity")); print("pipeline", d.get("pipeline"))
print("aop", d["aop"]["value"], d["aop"]["roofline"]["frac"], d["aop"].get("cpu_baseline"), d["aop"].get("vector"))
for k, x in d.get("extras", {}).items():
    print(k, x["value"], x["unit"], "e2e", x["e2e"]["value"], x["e2e"].get("pageable_host_value"), "roof", x["roofline"]["kernel"], round(x["roofline"]["frac"], 4), "cpu", x.get("cpu_baseline", {}).get("value"), "parity", {kk: x.get("parity", {}).get(kk) for kk in ("n", "max_rel", "n_over_1e_8", "n_fourier_mismatch", "pass")})
PY
