#!/usr/bin/env python
"""bench.py -- measures the AeroApps hot path on B200 (contract: one JSON line on rank 0).

  python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input.
`value` is whole-job throughput with inputs resident in HBM; `e2e` is the same metric through the
host-pointer C-ABI entry (pinned host buffers, H2D + D2H inside the timed region).  `roofline`
describes the dominant kernel; `cpu_baseline` is the CPU oracle timed on a bounded sample.
`--impl reference` times the reference algorithm's CPU restatement (oracle/) on all host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# every rank owns its own contiguous slab of the (virtual) swath: different synthetic pixels per rank
RANK_SEED = 97 * int(os.environ.get("RANK", "0"))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "_source": "fallback"}


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


# =================================================================================================
# workload: Stage A scalar AOP, BASELINE config 1 (c90: 48 600 columns x 72 levels, 20 tracers, 550 nm)
class AopScalarWorkload:
    name = "mie_aop_scalar_c90"
    metric = "Mie AOP voxels/sec (tau, ssa, g at 550 nm; 72 levels, 20 GOCART tracers, 7 species tables)"
    unit = "voxels/s"
    dtype = "f32"
    km, nobs, nch = 72, 48600, 1

    def __init__(self, nobs=None):
        from aeroapps_b200 import mietables as mt
        self.mt = mt
        if nobs:
            self.nobs = nobs
        self.tables = mt.synthetic_tables(nmom=0)
        self.chan_m = [550e-9]
        self.channels = np.array([550.0], np.float32)
        self.state = mt.synthetic_state(self.nobs, self.km, seed=20260101 + 1 + RANK_SEED)
        self.vnames = self.state["vnames"]
        self.nq = len(self.vnames)
        self.units_per_step = self.km * self.nobs * self.nch
        # algorithmic bytes per voxel: 4*(nq+1) in + 12*nch out
        self.bytes_per_unit = 4 * (self.nq + 1) + 12 * self.nch

    def config(self):
        return {"workload": self.name, "km": self.km, "nobs_per_gpu": self.nobs, "nq": self.nq, "nch": self.nch,
                "l2": "inputs (%.0f MB/step) larger than the 126 MB L2" % (self.units_per_step * 4 * (self.nq + 1) / 1e6)}

    # ---- GPU arm
    def setup_gpu(self, torch, dev):
        import ctypes as C
        from aeroapps_b200 import MieObs_
        from aeroapps_b200._lib import lib, check
        self.C, self.lib, self.check = C, lib(), check
        self.mie = MieObs_.ChemMie(self.chan_m, 0, self.tables)
        self.vn = self.mt.pad_names(self.vnames)
        st = self.state
        self.d_qm = torch.from_numpy(np.ascontiguousarray(st["qm"].T)).to(dev)      # memory == Fortran (km,nq,nobs)
        self.d_rh = torch.from_numpy(np.ascontiguousarray(st["rh"].T)).to(dev)
        nout = self.units_per_step
        self.d_out = [torch.empty(nout, dtype=torch.float32, device=dev) for _ in range(3)]
        # pinned host buffers for the e2e leg
        self.h_qm = torch.from_numpy(np.ascontiguousarray(st["qm"].T)).pin_memory()
        self.h_rh = torch.from_numpy(np.ascontiguousarray(st["rh"].T)).pin_memory()
        self.h_out = [torch.empty(nout, dtype=torch.float32).pin_memory() for _ in range(3)]
        self.h2d = self.h_qm.numel() * 4 + self.h_rh.numel() * 4
        self.d2h = 3 * nout * 4
        self.launches_per_step = 1

    def step_gpu(self, stream):
        C = self.C
        p = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_float))
        rc = C.c_int(0)
        self.check(self.lib.ab200_get_aop_scalar_dev(
            self.mie.handle, self.km, self.nobs, self.nch, self.nq, self.channels.ctypes.data_as(C.POINTER(C.c_float)),
            self.vn.ctypes.data_as(C.c_char_p), p(self.d_qm), p(self.d_rh), p(self.d_out[0]), p(self.d_out[1]),
            p(self.d_out[2]), C.c_void_p(stream), C.byref(rc)), "aop_scalar_dev")
        assert rc.value == 0

    def step_e2e(self):
        C = self.C
        p = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_float))
        rc = C.c_int(0)
        self.check(self.lib.ab200_get_aop_scalar(
            self.mie.handle, self.km, self.nobs, self.nch, self.nq, self.channels.ctypes.data_as(C.POINTER(C.c_float)),
            self.vn.ctypes.data_as(C.c_char_p), 0, p(self.h_qm), p(self.h_rh), p(self.h_out[0]), p(self.h_out[1]),
            p(self.h_out[2]), C.byref(rc)), "aop_scalar")
        assert rc.value == 0

    # ---- CPU arm (oracle = restatement of the reference Fortran; test infrastructure used as checker/baseline)
    def setup_cpu(self, sample_cols):
        from oracle import mie_oracle as orc
        self.orc = orc.OracleMie(self.chan_m, 0, self.tables)
        self.sample_cols = min(sample_cols, self.nobs)
        st = self.state
        self.c_qm = np.asfortranarray(st["qm"][:, :, :self.sample_cols])
        self.c_rh = np.asfortranarray(st["rh"][:, :self.sample_cols])

    def step_cpu(self, nthreads):
        from concurrent.futures import ThreadPoolExecutor
        n = self.sample_cols
        edges = np.linspace(0, n, nthreads + 1).astype(int)
        chunks = [(np.asfortranarray(self.c_qm[:, :, a:b]), np.asfortranarray(self.c_rh[:, a:b]))
                  for a, b in zip(edges[:-1], edges[1:]) if b > a]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(nthreads) as ex:
            list(ex.map(lambda c: self.orc.getaopscalar(self.channels, self.vnames, c[0], c[1]), chunks))
        dt = time.perf_counter() - t0
        return self.km * n * self.nch, dt

    def roofline(self, kernel_ms):
        pk = peaks()
        achieved = self.units_per_step * self.bytes_per_unit / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "r01b_rt_traffic.json")
        if os.path.exists(tp) and self.nobs == 48600:
            tj = json.load(open(tp)).get("k_aop_scalar")
            if tj:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        return {"bound": "hbm", "kernel": "k_aop_scalar_direct", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk["_source"],
                "bytes_per_unit": self.bytes_per_unit}


# =================================================================================================
# workload: Stage B vector VLIDORT (BASELINE configs[2]: nstokes=3, 12 streams, 72 layers, Greek moments from the
# Mie tables, nMom=300), processed in batches; the Mie-AOP stage that feeds it is timed as `aop` extra info.
class RtVectorWorkload:
    name = "vlidort_vector_n12_l72_lambert"
    metric = "VLIDORT pixels/sec (vector nstokes=3, 12 streams, 72 layers, nMom=300, FO + delta-M, Lambertian surface)"
    unit = "pixels/s"
    dtype = "f64"
    km, nobs, nch, N, S, nmom = 72, 2048, 1, 12, 3, 300

    def __init__(self, nobs=None):
        from aeroapps_b200 import mietables as mt
        self.mt = mt
        if nobs:
            self.nobs = nobs
        self.units_per_step = self.nobs * self.nch
        self.plane_parallel = False

    def config(self):
        n = self.S * self.N
        return {"workload": self.name, "km": self.km, "pixels_per_step_per_gpu": self.nobs, "nstokes": self.S, "nstreams": self.N,
                "nMom": self.nmom, "plane_parallel": self.plane_parallel,
                "l2": "per-step working set (%.1f GB of layer solutions and BVP factors) far larger than the 126 MB L2"
                      % (self.nobs * self.km * (2 * n * n + 20 * n) * 8 / 1e9)}

    def _scene_gpu(self):
        """synthetic pixels built with the PRODUCT's own Stage A (CUDA): Mie AOP vector -> edge vars -> Rayleigh"""
        from aeroapps_b200 import MieObs_, VLIDORT_POLAR_ as VP
        mt = self.mt
        tables = mt.synthetic_tables(nmom=self.nmom)
        st = mt.synthetic_state(self.nobs, self.km, seed=20260101 + 3 + RANK_SEED)
        mie = MieObs_.ChemMie([550e-9], self.nmom, tables)
        t0 = time.perf_counter()
        tau, ssa, g, pmom, rc = mie.getaopvector([550.0], st["vnames"], 0, st["qm"], st["rh"], self.nmom, 6)
        self.aop_e2e_voxels_per_s = self.km * self.nobs / (time.perf_counter() - t0)
        assert rc == 0
        self.aop_vector = self._time_aop_vector(mie, st)
        pe, ze, te = MieObs_.getedgevars(st["airdens"], st["delp"], 1.0)
        self.mie, self.state = mie, st                       # kept for the fused-pipeline leg
        pmom = pmom.astype(np.float64); pmom[..., 1] *= -1.0; pmom[..., 3] *= -1.0
        F = lambda a: np.asfortranarray(a, dtype=np.float64)
        ch = np.array([550.0])
        ROT, depol, rc = VP.rot_calc(ch, F(pe), F(ze), F(te), 1e15, 0)
        rng = np.random.default_rng(20260101 + 4 + RANK_SEED)
        n = self.nobs
        return dict(channels=ch, ROT=F(ROT), depol=depol, alpha=np.zeros_like(F(ROT)), tau=F(tau), ssa=F(ssa), pmom=F(pmom),
                    pe=F(pe), he=F(ze), te=F(te), albedo=F(rng.uniform(0.02, 0.4, (n, 1))), sza=rng.uniform(0, 75, n),
                    raa=rng.uniform(0, 180, n), vza=rng.uniform(0, 65, n), flux=np.ones((1, n), order="F"))

    def _time_aop_vector(self, mie, st):
        """k_aop_vector (getAOPvector with the nMom x 6 phase-matrix moments) on device-resident inputs: voxels/s and the
        HBM bandwidth its algorithmic bytes (4 (nq+1) + 4 nch (3 + nMom nPol) per voxel) amount to"""
        import ctypes as C
        import torch
        from aeroapps_b200._lib import lib, check
        dev = torch.device("cuda", torch.cuda.current_device())
        nq, nvox = len(st["vnames"]), self.km * self.nobs
        d_qm = torch.from_numpy(np.ascontiguousarray(st["qm"].T)).to(dev); d_rh = torch.from_numpy(np.ascontiguousarray(st["rh"].T)).to(dev)
        outs = [torch.empty(nvox, dtype=torch.float32, device=dev) for _ in range(3)]
        d_pm = torch.empty(nvox * self.nmom * 6, dtype=torch.float32, device=dev)
        ch = np.array([550.0], np.float32); vn = self.mt.pad_names(st["vnames"])
        p = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_float))
        rc = C.c_int(0)
        stream = torch.cuda.current_stream().cuda_stream

        def run():
            check(lib().ab200_get_aop_vector_dev(mie.handle, self.km, self.nobs, 1, nq, ch.ctypes.data_as(C.POINTER(C.c_float)),
                                                 vn.ctypes.data_as(C.c_char_p), p(d_qm), p(d_rh), self.nmom, 6, p(outs[0]), p(outs[1]),
                                                 p(outs[2]), p(d_pm), C.c_void_p(stream), C.byref(rc)), "aop_vector_dev")
        for _ in range(3):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(5):
            run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        bpu = 4 * (nq + 1) + 4 * (3 + self.nmom * 6)
        gbs = nvox * bpu / (ms * 1e-3) / 1e9
        pk = peaks()
        return {"kernel": "k_aop_vector", "voxels_per_s": nvox / (ms * 1e-3), "ms": ms, "bytes_per_unit": bpu, "achieved_gbs": gbs,
                "frac_of_hbm_peak": gbs / pk["hbm_gbs"], "voxels": nvox, "nMom": self.nmom, "nPol": 6}

    def setup_gpu(self, torch, dev):
        import ctypes as C
        from aeroapps_b200._lib import lib, check
        self.C, self.lib, self.check, self.torch = C, lib(), check, torch
        sc = self.scene = self._scene_gpu()
        keys = ["ROT", "depol", "alpha", "tau", "ssa", "pmom", "he", "albedo", "sza", "raa", "vza", "flux"]
        flat = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F")))
        self.h = {k: flat(sc[k]).pin_memory() for k in keys + ["pe", "te", "channels"]}
        self.d = {k: self.h[k].to(dev) for k in keys}
        self.d_out = [torch.empty(self.nobs * self.nch, dtype=torch.float64, device=dev) for _ in range(4)]
        self.h_out = [torch.empty(self.nobs * self.nch, dtype=torch.float64).pin_memory() for _ in range(4)]
        self.h2d = sum(self.h[k].numel() * 8 for k in keys + ["pe", "te"])
        self.d2h = 4 * self.nobs * self.nch * 8
        self.lib.ab200_vlidort_stats(None, None, None, None, None, None, 1)

    def _p(self, t):
        return self.C.cast(t.data_ptr(), self.C.POINTER(self.C.c_double))

    def step_gpu(self, stream):
        C, d, p = self.C, self.d, self._p
        rc = C.c_int(0)
        self.check(self.lib.ab200_vlidort_vector_lambert_dev(
            self.km, self.nch, self.nobs, self.N, int(self.plane_parallel), self.nmom, 6, p(d["ROT"]), p(d["depol"]), p(d["alpha"]),
            p(d["tau"]), p(d["ssa"]), p(d["pmom"]), p(d["he"]), p(d["albedo"]), p(d["sza"]), p(d["raa"]), p(d["vza"]), p(d["flux"]),
            C.c_double(1e15), *[p(o) for o in self.d_out], C.c_void_p(stream), C.byref(rc)), "vector_lambert_dev")
        assert rc.value == 0

    def step_e2e(self):
        C, h, p = self.C, self.h, self._p
        rc = C.c_int(0)
        self.check(self.lib.ab200_vlidort_vector_lambert(
            self.km, self.nch, self.nobs, p(h["channels"]), self.N, int(self.plane_parallel), self.nmom, 6, p(h["ROT"]), p(h["depol"]),
            p(h["alpha"]), p(h["tau"]), p(h["ssa"]), p(h["pmom"]), p(h["pe"]), p(h["he"]), p(h["te"]), p(h["albedo"]), p(h["sza"]),
            p(h["raa"]), p(h["vza"]), p(h["flux"]), C.c_double(1e15), 0, *[p(o) for o in self.h_out], C.byref(rc)), "vector_lambert")
        assert rc.value == 0

    def stats(self, reset=False):
        C = self.C
        a, b, c = C.c_double(), C.c_double(), C.c_double(); d, e, f = C.c_long(), C.c_long(), C.c_long()
        fe, fb, nl = C.c_double(), C.c_double(), C.c_long()
        self.lib.ab200_vlidort_stats2(C.byref(fe), C.byref(fb), C.byref(nl))
        self.lib.ab200_vlidort_stats(C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e), C.byref(f), int(reset))
        return dict(ms_setup=a.value, ms_eigen=b.value, ms_bvp=c.value, n_eigen=d.value, n_bvp=e.value, npix=f.value,
                    flops_eigen=fe.value, flops_bvp=fb.value, n_launches=nl.value)

    # ---- CPU arm: the oracle restatement of the reference at the wrapper level
    def setup_cpu(self, sample_cols):
        from oracle import vlidort_glue as vg
        self.vg = vg
        cores = os.cpu_count() or 1
        self.sample_cols = max(cores, min(sample_cols, 16 * cores, self.nobs))
        if getattr(self, "scene", None) is None:          # --impl reference: no GPU -> inputs from the CPU oracle of Stage A
            sc = vg.synthetic_scene(self.sample_cols, nmom=self.nmom, seed=20260101 + 3)
            sc["flux"] = np.ones((1, self.sample_cols), order="F"); sc["alpha"] = np.zeros_like(sc["ROT"])
            self.cpu_scene = sc
        else:
            s = self.scene; n = self.sample_cols
            self.cpu_scene = dict(ROT=s["ROT"][:, :n], depol=s["depol"], alpha=s["alpha"][:, :n], tau=s["tau"][:, :, :n], ssa=s["ssa"][:, :, :n],
                                  pmom=s["pmom"][:, :, :n], he=s["he"][:, :n], albedo=s["albedo"][:n], sza=s["sza"][:n], raa=s["raa"][:n],
                                  vza=s["vza"][:n], flux=s["flux"][:, :n])

    def step_cpu(self, nthreads):
        from concurrent.futures import ThreadPoolExecutor
        s, n, vg = self.cpu_scene, self.sample_cols, self.vg
        edges = np.linspace(0, n, nthreads + 1).astype(int)
        rng = [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(nthreads) as ex:
            list(ex.map(lambda ab: vg.vlidort_run(3, self.N, self.plane_parallel, s["ROT"], s["depol"], s["tau"], s["ssa"], s["pmom"], s["he"],
                                                  s["sza"], s["raa"], s["vza"], albedo=s["albedo"], alpha=s["alpha"], flux_factor=s["flux"],
                                                  obs_range=ab), rng))
        return n * self.nch, time.perf_counter() - t0

    def roofline(self, step_ms):
        st = self.stats()
        fp = os.path.join(ROOT, "profiles", "r01_fp64_peak.json")
        peak = json.load(open(fp))["dfma_tflops"] if os.path.exists(fp) else 37.0
        dom = "k_vl_eigen" if st["ms_eigen"] >= st["ms_bvp"] else "k_vl_bvp"
        ms = st["ms_eigen"] if dom == "k_vl_eigen" else st["ms_bvp"]
        fl = st["flops_eigen"] if dom == "k_vl_eigen" else st["flops_bvp"]
        ach = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        tot = (st["flops_eigen"] + st["flops_bvp"]) / ((st["ms_eigen"] + st["ms_bvp"] + st["ms_setup"]) * 1e-3) / 1e12
        traffic, tnote = None, None
        tp = os.path.join(ROOT, "profiles", "r01b_rt_traffic.json")
        if os.path.exists(tp) and self.nobs == 2048:
            tj = json.load(open(tp)).get(dom)
            if tj:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                tnote = "dram read+write bytes of the m=0 launch (2048 pixels) from the committed ncu --set full capture"
        return {"bound": "fp64", "kernel": dom, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "traffic_note": tnote,
                "peak_source": "measured DFMA microbenchmark on this pool (tools/fp64_peak.cu, profiles/r01_fp64_peak.json); "
                               "MEASURED_PEAKS.json holds no FP64 figure",
                "whole_rt_tflops": tot, "kernel_ms_share": {"setup": st["ms_setup"], "eigen": st["ms_eigen"], "bvp": st["ms_bvp"]},
                "fourier_terms_per_pixel": st["n_bvp"] / max(st["npix"], 1)}


class RtVectorCxWorkload(RtVectorWorkload):
    """BASELINE.json configs[2]: vector VLIDORT, 12 streams, 72 layers, Greek moments from the Mie tables, Cox-Munk ocean
    BRDF (f2py VECTOR_CX), leo_vlidort-style swath of synthetic pixels."""
    name = "vlidort_vector_n12_l72_coxmunk"
    metric = "VLIDORT pixels/sec (vector nstokes=3, 12 streams, 72 layers, nMom=300, FO + delta-M, Cox-Munk BRDF)"

    def _scene_gpu(self):
        sc = super()._scene_gpu()
        rng = np.random.default_rng(20260101 + 5 + RANK_SEED)
        n = self.nobs
        sc.update(u10m=rng.normal(0, 5, n), v10m=rng.normal(0, 5, n), mr=np.array([1.334]))
        return sc

    def setup_gpu(self, torch, dev):
        super().setup_gpu(torch, dev)
        flat = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F")))
        for k in ("u10m", "v10m", "mr"):
            self.h[k] = flat(self.scene[k]).pin_memory(); self.d[k] = self.h[k].to(dev)
        # radiance, reflectance, BR, Q, U (+ BR_Q, BR_U on the host path)
        self.d_out = [torch.empty(self.nobs * self.nch, dtype=torch.float64, device=dev) for _ in range(5)]
        self.h_out = [torch.empty(self.nobs * self.nch, dtype=torch.float64).pin_memory() for _ in range(7)]
        keys = ["ROT", "depol", "tau", "ssa", "pmom", "pe", "he", "te", "u10m", "v10m", "mr", "sza", "raa", "vza"]
        self.h2d = sum(self.h[k].numel() * 8 for k in keys)
        self.d2h = 7 * self.nobs * self.nch * 8

    def step_gpu(self, stream):
        C, d, p = self.C, self.d, self._p
        rc = C.c_int(0)
        self.check(self.lib.ab200_vlidort_vector_cx_dev(
            self.km, self.nch, self.nobs, self.N, int(self.plane_parallel), self.nmom, 6, p(d["ROT"]), p(d["depol"]), p(d["tau"]),
            p(d["ssa"]), p(d["pmom"]), p(d["he"]), p(d["u10m"]), p(d["v10m"]), p(d["mr"]), p(d["sza"]), p(d["raa"]), p(d["vza"]),
            C.c_double(1e15), *[p(o) for o in self.d_out], C.c_void_p(stream), C.byref(rc)), "vector_cx_dev")
        assert rc.value == 0

    def step_e2e(self):
        C, h, p = self.C, self.h, self._p
        rc = C.c_int(0)
        self.check(self.lib.ab200_vlidort_vector_cx(
            self.km, self.nch, self.nobs, p(h["channels"]), self.N, int(self.plane_parallel), self.nmom, 6, p(h["ROT"]), p(h["depol"]),
            p(h["tau"]), p(h["ssa"]), p(h["pmom"]), p(h["pe"]), p(h["he"]), p(h["te"]), p(h["u10m"]), p(h["v10m"]), p(h["mr"]),
            p(h["sza"]), p(h["raa"]), p(h["vza"]), C.c_double(1e15), 0, *[p(o) for o in self.h_out], C.byref(rc)), "vector_cx")
        assert rc.value == 0

    def pipeline_leg(self, torch, steps=3):
        """the same pixels through ab200_pipeline_vector_cx (getEdgeVars -> ROT_CALC -> getAOPvector -> VECTOR_CX in one call):
        host inputs are the model state, every intermediate stays in HBM.  Returns pixels/s and the bytes that cross PCIe."""
        from aeroapps_b200 import pipeline
        st, sc = self.state, self.scene
        a = (self.mie, [550.0], st["vnames"], st["qm"], st["rh"], st["airdens"], st["delp"], 1.0, self.N, self.plane_parallel,
             self.nmom, 6, sc["u10m"], sc["v10m"], sc["mr"], sc["sza"], sc["raa"], sc["vza"], 1e15, 0)
        out = pipeline.vector_cx_pipeline(*a)
        assert out[-1] == 0
        same = bool(np.array_equal(out[0].ravel(), self.h_out[0].numpy()) and np.array_equal(out[3].ravel(), self.h_out[3].numpy()))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            pipeline.vector_cx_pipeline(*a)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        h2d = 4 * (st["qm"].size + st["rh"].size + st["airdens"].size + st["delp"].size) + 8 * 5 * self.nobs
        return {"value": self.units_per_step / dt, "unit": self.unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 5 * 8 * self.nobs,
                "bit_identical_to_separate_calls": same,
                "entry": "ab200_pipeline_vector_cx (fused getEdgeVars -> ROT_CALC -> getAOPvector -> VECTOR_CX, per-rank wall clock)"}

    def setup_cpu(self, sample_cols):
        had_scene = getattr(self, "scene", None) is not None
        super().setup_cpu(sample_cols)
        n = self.sample_cols
        if had_scene:
            self.cpu_scene.update(u10m=self.scene["u10m"][:n], v10m=self.scene["v10m"][:n], mr=self.scene["mr"])

    def step_cpu(self, nthreads):
        from concurrent.futures import ThreadPoolExecutor
        s, n, vg = self.cpu_scene, self.sample_cols, self.vg
        edges = np.linspace(0, n, nthreads + 1).astype(int)
        rng = [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(nthreads) as ex:
            list(ex.map(lambda ab: vg.vlidort_run_surface("cx", 3, self.N, self.plane_parallel, s["ROT"], s["depol"], s["tau"], s["ssa"],
                                                          s["pmom"], s["he"], s["sza"], s["raa"], s["vza"], s["u10m"], s["v10m"], s["mr"],
                                                          obs_range=ab), rng))
        return n * self.nch, time.perf_counter() - t0


class RtScalarWorkload(RtVectorWorkload):
    """BASELINE.json configs[1]: scalar VLIDORT (nstokes = 1), 8 streams, 72 layers, Lambertian surface, nMom = 64 (not the headline
    workload; `python bench.py --workload rt_scalar`).  Only the reference-style host entry exists for the scalar form, so the
    device-resident leg runs the same call: `value` includes the PCIe copies here and equals `e2e` up to noise."""
    name = "vlidort_scalar_n8_l72_lambert"
    metric = "VLIDORT pixels/sec (scalar nstokes=1, 8 streams, 72 layers, nMom=64, FO + delta-M, Lambertian surface)"
    km, nobs, nch, N, S, nmom = 72, 16384, 1, 8, 1, 64

    def setup_gpu(self, torch, dev):
        super().setup_gpu(torch, dev)
        g = np.zeros(self.km * self.nobs)                      # the scalar wrapper takes g but does not use it
        self.h["g"] = torch.from_numpy(g).pin_memory()
        self.h_out = [torch.empty(self.nobs * self.nch, dtype=torch.float64).pin_memory() for _ in range(2)]
        keys = ["ROT", "depol", "tau", "ssa", "g", "pmom", "pe", "he", "te", "albedo", "sza", "raa", "vza"]
        self.h2d = sum(self.h[k].numel() * 8 for k in keys)
        self.d2h = 2 * self.nobs * self.nch * 8

    def step_e2e(self):
        C, h, p = self.C, self.h, self._p
        rc = C.c_int(0)
        self.check(self.lib.ab200_vlidort_scalar_lambert(
            self.km, self.nch, self.nobs, p(h["channels"]), self.N, int(self.plane_parallel), self.nmom, 6, p(h["ROT"]), p(h["depol"]),
            p(h["tau"]), p(h["ssa"]), p(h["g"]), p(h["pmom"]), p(h["pe"]), p(h["he"]), p(h["te"]), p(h["albedo"]), p(h["sza"]),
            p(h["raa"]), p(h["vza"]), C.c_double(1e15), 0, p(self.h_out[0]), p(self.h_out[1]), C.byref(rc)), "scalar_lambert")
        assert rc.value == 0

    def step_gpu(self, stream):
        self.step_e2e()

    def step_cpu(self, nthreads):
        from concurrent.futures import ThreadPoolExecutor
        s, n, vg = self.cpu_scene, self.sample_cols, self.vg
        edges = np.linspace(0, n, nthreads + 1).astype(int)
        rng = [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]
        t0 = time.perf_counter()
        with ThreadPoolExecutor(nthreads) as ex:
            list(ex.map(lambda ab: vg.vlidort_run(1, self.N, self.plane_parallel, s["ROT"], s["depol"], s["tau"], s["ssa"], s["pmom"], s["he"],
                                                  s["sza"], s["raa"], s["vza"], albedo=s["albedo"], scalar_glue=True, obs_range=ab), rng))
        return n * self.nch, time.perf_counter() - t0


WORKLOADS = {"aop_scalar": AopScalarWorkload, "rt_vector": RtVectorWorkload, "rt_vector_cx": RtVectorCxWorkload, "rt_scalar": RtScalarWorkload}
DEFAULT_WORKLOAD = "rt_vector_cx"


def run_reference(args, wl):
    """--impl reference: the reference algorithm's CPU implementation (oracle restatement; the Fortran
    itself cannot be built here: no Fortran compiler, MAPL/ESMF un-vendored) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    wl.setup_cpu(sample_cols=args.cpu_sample)
    for _ in range(args.warmup):
        wl.step_cpu(cores)
    units = 0; secs = 0.0
    for _ in range(args.steps):
        u, dt = wl.step_cpu(cores)
        units += u; secs += dt
    v = units / secs
    sample = f"{wl.sample_cols} of {wl.nobs} columns per step"
    print(json.dumps({
        "impl": "reference", "metric": wl.metric, "value": v, "unit": wl.unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": wl.config(),
        "cpu_baseline": {"value": v, "unit": wl.unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_gpu(wl, args, torch, dev, dist, rank, world, local):
    """W warm-up + K timed steps of one workload (device-resident inputs), then the e2e leg; returns the JSON dict"""
    wl.setup_gpu(torch, dev)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        wl.step_gpu(stream)
    barrier()
    if hasattr(wl, "stats"):
        wl.stats(reset=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local) as cs:
        barrier()
        ev[0].record()
        for i in range(args.steps):
            wl.step_gpu(stream)
            ev[i + 1].record()
        barrier()
        total_ms = ev[0].elapsed_time(ev[-1])
        roof = wl.roofline(total_ms / args.steps) if rank == 0 else None
        launches = wl.stats()["n_launches"] if hasattr(wl, "stats") else wl.launches_per_step * args.steps
        # keep the sampler alive long enough for at least a few samples on very short runs
        if total_ms < 1500:
            t_end = time.time() + 1.5
            while time.time() < t_end:
                wl.step_gpu(stream)
            torch.cuda.synchronize()
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = wl.units_per_step * world * args.steps / (total_ms * 1e-3)

    # e2e leg: host (pinned) buffers through the host-pointer C-ABI call, copies inside the timed region
    for _ in range(2):
        wl.step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, min(args.steps, 10))
    e0.record()
    for _ in range(e2e_steps):
        wl.step_e2e()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = wl.units_per_step * world * e2e_steps / (float(t.item()) * 1e-3)
    if rank != 0:
        return None
    return {
        "metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": wl.config(), "roofline": roof,
        "e2e": {"value": e2e_value, "unit": wl.unit, "h2d_bytes_per_step": wl.h2d, "d2h_bytes_per_step": wl.d2h},
        "gpu_launches": int(launches), "clocks": cs.summary(),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--nobs", type=int, default=0, help="override columns/pixels per GPU per step (debug)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="columns/pixels in the bounded CPU sample (0 = workload default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-aop", action="store_true", help="skip the extra Mie-AOP measurement attached to the RT line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if not args.cpu_sample:
        args.cpu_sample = 48600 if args.workload == "aop_scalar" else 10 ** 9

    wl = WORKLOADS[args.workload](nobs=args.nobs or None)
    if args.impl == "reference":
        return run_reference(args, wl)

    import torch
    import __graft_entry__ as g
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: NCCL's own banner ("NCCL version ...", printed to stdout when the
        # environment sets NCCL_DEBUG) goes to stderr while the communicator is created
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    if rank == 0:
        g.build()
    if dist:
        dist.barrier()
    out = run_gpu(wl, args, torch, dev, dist, rank, world, local)
    if dist and hasattr(wl, "d_out"):
        # the path's only exchange: the final gather of the per-rank slabs (outside the timed steps)
        from aeroapps_b200 import swath
        loc = wl.d_out[0].detach().cpu().numpy()
        full = swath.gather_swath(loc, loc.shape[0] * world, axis=0, dist=dist, device=dev)
        if rank == 0:
            out["config"]["gathered_outputs"] = int(full.shape[0])
            out["config"]["parallelism"] = f"{world} ranks x contiguous slabs, no data-path collective, final all_gather over NCCL"
    if hasattr(wl, "pipeline_leg"):
        pl = wl.pipeline_leg(torch)
        if rank == 0:
            out["pipeline"] = pl
    if args.workload.startswith("rt_vector") and not args.no_aop:
        # the second half of the BASELINE metric: Mie AOP voxels/s (BASELINE configs[0], c90 scalar AOPs)
        import copy
        a2 = copy.copy(args); a2.steps = 20; a2.warmup = 5
        aop = run_gpu(AopScalarWorkload(), a2, torch, dev, dist, rank, world, local)
        if rank == 0:
            out["aop"] = {k: aop[k] for k in ("metric", "value", "unit", "ms_per_step", "roofline", "e2e", "config")}
            out["aop"]["vector_e2e_voxels_per_s_nmom300"] = getattr(wl, "aop_e2e_voxels_per_s", None)
            out["aop"]["vector"] = getattr(wl, "aop_vector", None)
    if rank == 0:
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            wl.setup_cpu(sample_cols=args.cpu_sample)
            u, dt = wl.step_cpu(cores)
            out["cpu_baseline"] = {"value": u / dt, "unit": wl.unit, "cores": cores, "kind": "port",
                                   "sample": f"{wl.sample_cols} of {wl.nobs} {wl.unit.split('/')[0]} of the same workload, 1 pass, "
                                             f"{cores} threads (oracle/: C/C++ restatement of the reference Fortran)"}
        print(json.dumps(out))
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
