"""K3 phase timers (debug build: make -C aeroapps_b200/csrc clean all EXTRA=-DAB200_BVP_TIMING).
Runs the bench workload for a few steps and prints the share of each phase of bvp_pixel as seen by thread 0."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

NAMES = ["init", "load rows", "first panel+urows", "lookahead+sync", "panel", "trailing join+sync", "urows+sync", "spill/shift",
         "bwd rhs", "tri solve", "source terms", "v4 publish", "v4 barrier B", "13", "14", "15"]
# v4 (register-resident elimination) reuses the slots: 1 new rows, 2 first pivot, 3 multiplier+spill, 4 phase 0 (rank-1 update or
# next column + candidate), 5 barrier A, 6 select + rank-1 update of the column owners, 11 publish, 12 barrier B, 7 layer end
wl = bench.WORKLOADS[bench.DEFAULT_WORKLOAD](int(sys.argv[1]) if len(sys.argv) > 1 else None)
dev = torch.device("cuda:0")
wl.setup_gpu(torch, dev)
st = torch.cuda.current_stream().cuda_stream
wl.step_gpu(st); torch.cuda.synchronize()
buf = (C.c_ulonglong * 16)()
wl.lib.ab200_vlidort_debug_timing(buf)
wl.stats(reset=True)
for _ in range(2):
    wl.step_gpu(st)
torch.cuda.synchronize()
wl.lib.ab200_vlidort_debug_timing(buf)
s = wl.stats()
tot = sum(buf)
out = {"ms_bvp": s["ms_bvp"], "ms_eigen": s["ms_eigen"], "n_bvp": s["n_bvp"], "cycles_per_pixel_term": tot / max(s["n_bvp"], 1),
       "phases": {n: round(buf[i] / tot, 4) for i, n in enumerate(NAMES)}}
print(json.dumps(out, indent=1))
