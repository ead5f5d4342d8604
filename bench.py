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

    def step_pageable(self):
        """the same call with ordinary (pageable) numpy arrays: what the f2py module's caller has"""
        C = self.C
        if not hasattr(self, "p_qm"):
            st = self.state
            self.p_qm = np.ascontiguousarray(st["qm"].T).copy(); self.p_rh = np.ascontiguousarray(st["rh"].T).copy()
            self.p_out = [np.empty(self.units_per_step, np.float32) for _ in range(3)]
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
        rc = C.c_int(0)
        self.check(self.lib.ab200_get_aop_scalar(
            self.mie.handle, self.km, self.nobs, self.nch, self.nq, self.channels.ctypes.data_as(C.POINTER(C.c_float)),
            self.vn.ctypes.data_as(C.c_char_p), 0, p(self.p_qm), p(self.p_rh), p(self.p_out[0]), p(self.p_out[1]),
            p(self.p_out[2]), C.byref(rc)), "aop_scalar")
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
# Stage B workloads (BASELINE.json configs[1..4]): one generic class, parameterised by nstokes / streams / surface / channels / views.
# The scene is drawn from fixed seeds (column state, angles, surface); Stage A (Mie AOP vector -> edge variables -> Rayleigh) runs on the
# GPU through the PRODUCT for the GPU arm and through the ORACLE for the CPU arms: both are bit-identical (tests/test_mie_gpu.py), so
# every arm sees the same pixels.
class RtWorkload:
    unit = "pixels/s"
    dtype = "f64"
    km = 72
    nch_nm = (550.0,)
    ngeom = 1
    plane_parallel = False
    seed = 20260101 + 3
    parity_pixels = 256          # pixels of the bench scene compared with the oracle (cpu_baseline leg)

    def __init__(self, nobs=None):
        from aeroapps_b200 import mietables as mt
        self.mt = mt
        if nobs:
            self.nobs = nobs
        self.nch = len(self.nch_nm)
        self.units_per_step = self.nobs * self.nch * self.ngeom
        self.scene = None

    def config(self):
        n = self.S * self.N
        c = {"workload": self.name, "km": self.km, "observations_per_step_per_gpu": self.nobs, "channels_nm": list(self.nch_nm),
             "views_per_observation": self.ngeom, "units_per_step_per_gpu": self.units_per_step, "nstokes": self.S, "nstreams": self.N,
             "nMom": self.nmom, "surface": self.kind, "plane_parallel": self.plane_parallel,
             "l2": "per-step working set (%.1f GB of layer solutions and interaction operators) far larger than the 126 MB L2"
                   % (self.nobs * self.nch * self.km * (6 * n * n + 20 * n) * 8 / 1e9)}
        return c

    # ---- scene ------------------------------------------------------------------------------------------------------------------
    def draw(self):
        """column state + geometry + surface of all nobs observations (host, small); the same numbers whoever asks"""
        n, G = self.nobs, self.ngeom
        st = self.mt.synthetic_state(n, self.km, seed=self.seed + RANK_SEED)
        rng = np.random.default_rng(self.seed + 1 + RANK_SEED)
        d = dict(state=st, sza=rng.uniform(0, 75, n), raa=rng.uniform(0, 180, (n, G)), vza=rng.uniform(0, 65, (n, G)),
                 albedo=rng.uniform(0.02, 0.4, (n, self.nch)), u10m=rng.normal(0, 5, n), v10m=rng.normal(0, 5, n),
                 mr=np.full(self.nch, 1.334),
                 kernel_wt=np.stack([rng.uniform(0.02, 0.4, (self.nch, n)), rng.uniform(0.0, 0.08, (self.nch, n)),
                                     rng.uniform(0.0, 0.2, (self.nch, n))]),
                 param=np.stack([np.full((self.nch, n), 2.0), np.full((self.nch, n), 1.0)]))
        return d

    def build_scene(self, provider, cols=None):
        """Fortran float64 arrays in the drop-in layouts for observations [0, cols); provider 'gpu' (product Stage A) or 'oracle'"""
        d = self.draw()
        n = self.nobs if cols is None else min(cols, self.nobs)
        st = d["state"]
        qm, rh, ad, dp = (np.asfortranarray(st[k][..., :n]) for k in ("qm", "rh", "airdens", "delp"))
        tables = self.mt.synthetic_tables(nmom=self.nmom)
        chan_m = [c * 1e-9 for c in self.nch_nm]
        F = lambda a: np.asfortranarray(a, dtype=np.float64)
        if provider == "gpu":
            from aeroapps_b200 import MieObs_, VLIDORT_POLAR_ as VP
            mie = MieObs_.ChemMie(chan_m, self.nmom, tables)
            t0 = time.perf_counter()
            tau, ssa, g, pmom, rc = mie.getaopvector(list(self.nch_nm), st["vnames"], 0, qm, rh, self.nmom, 6)
            self.aop_e2e_voxels_per_s = self.km * n * self.nch / (time.perf_counter() - t0)
            assert rc == 0
            pe, ze, te = MieObs_.getedgevars(ad, dp, 1.0)
            ROT, depol, rc = VP.rot_calc(np.array(self.nch_nm), F(pe), F(ze), F(te), 1e15, 0)
            self.mie, self.state_n = mie, dict(qm=qm, rh=rh, airdens=ad, delp=dp, vnames=st["vnames"])
        else:
            from oracle import mie_oracle as mo, vlidort_glue as vg
            om = mo.OracleMie(chan_m, self.nmom, tables)
            tau, ssa, g, pmom, rc = om.getaopvector(list(self.nch_nm), st["vnames"], qm, rh, self.nmom, 6)
            assert rc == 0
            pe, ze, te = mo.getedgevars(ad, dp, 1.0)
            ROT, depol = vg.rot_calc(self.nch_nm, pe, ze, te)
        pmom = pmom.astype(np.float64); pmom[..., 1] *= -1.0; pmom[..., 3] *= -1.0     # caller-side sign flip of P12, P34
        sc = dict(channels=np.array(self.nch_nm), ROT=F(ROT), depol=np.asarray(depol, dtype=np.float64), alpha=np.zeros_like(F(ROT)),
                  tau=F(tau), ssa=F(ssa), g=F(g), pmom=F(pmom), pe=F(pe), he=F(ze), te=F(te), albedo=F(d["albedo"][:n]),
                  sza=d["sza"][:n].copy(), raa=F(d["raa"][:n]), vza=F(d["vza"][:n]), flux=np.ones((self.nch, n), order="F"),
                  u10m=d["u10m"][:n].copy(), v10m=d["v10m"][:n].copy(), mr=d["mr"], kernel_wt=F(d["kernel_wt"][..., :n]),
                  param=F(d["param"][..., :n]))
        if self.ngeom == 1:
            sc["raa"] = sc["raa"][:, 0].copy(); sc["vza"] = sc["vza"][:, 0].copy()
        return sc

    # ---- the reference-style host entry (ctypes mirror of the f2py module) with whatever host arrays it is given ---------------------
    def call_host(self, sc):
        from aeroapps_b200 import VLIDORT_POLAR_ as VP
        a = (sc["channels"], self.N, self.plane_parallel, sc["ROT"], sc["depol"])
        b = (sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"], sc["te"])
        geo = (sc["sza"], sc["raa"], sc["vza"])
        G = self.ngeom
        szg = np.asfortranarray(np.repeat(sc["sza"][:, None], G, axis=1)) if G > 1 else None
        if self.kind == "lambert" and self.S == 1:
            I, R, rc = VP.scalar_lambert(*a, sc["tau"], sc["ssa"], sc["g"], sc["pmom"], sc["pe"], sc["he"], sc["te"], sc["albedo"], *geo, 1e15, 0)
            Q = U = None
        elif self.kind == "lambert" and G == 1:
            I, R, Q, U, rc = VP.vector_lambert(*a, sc["alpha"], *b, sc["albedo"], *geo, sc["flux"], 1e15, 0)
        elif self.kind == "lambert":
            I, R, Q, U, rc = VP.vector_lambert_multigeom(*a, sc["alpha"], *b, sc["albedo"], szg, sc["raa"], sc["vza"], sc["flux"], 1e15, 0)
        elif self.kind == "cx" and G == 1:
            I, R, BR, Q, U, _, _, rc = VP.vector_cx(*a, *b, sc["u10m"], sc["v10m"], sc["mr"], *geo, 1e15, 0)
        elif self.kind == "cx":
            I, R, BR, Q, U, _, _, rc = VP.vector_cx_multigeom(*a, *b, sc["u10m"], sc["v10m"], sc["mr"], szg, sc["raa"], sc["vza"], 1e15, 0)
        elif self.kind == "modis":
            I, R, BR, Q, U, _, _, rc = VP.vector_brdf_modis(*a, *b, sc["kernel_wt"], sc["param"], *geo, 1e15, 0)
        else:
            raise ValueError(self.kind)
        assert rc == 0
        nf, st = VP.last_diag()
        return dict(radiance=I, reflectance=R, Q=Q, U=U, nfourier=nf, status=st)

    BIG = ("ROT", "alpha", "tau", "ssa", "g", "pmom", "pe", "he", "te", "albedo", "raa", "vza", "flux", "kernel_wt", "param")

    def setup_gpu(self, torch, dev):
        import ctypes as C
        from aeroapps_b200._lib import lib, check
        self.C, self.lib, self.check, self.torch = C, lib(), check, torch
        sc = self.scene = self.build_scene("gpu")
        # pinned copies of the large arrays: numpy views on page-locked torch storage, in the same Fortran layouts
        self.pinned, self._keep = dict(sc), []
        for k in self.BIG:
            a = sc[k]
            t = torch.from_numpy(np.ascontiguousarray(a.ravel(order="F"))).pin_memory()
            self._keep.append(t)
            self.pinned[k] = t.numpy().reshape(a.shape, order="F")
        used = {"lambert": ("ROT", "depol", "alpha", "tau", "ssa", "pmom", "pe", "he", "te", "albedo", "sza", "raa", "vza", "flux"),
                "cx": ("ROT", "depol", "tau", "ssa", "pmom", "pe", "he", "te", "u10m", "v10m", "mr", "sza", "raa", "vza"),
                "modis": ("ROT", "depol", "tau", "ssa", "pmom", "pe", "he", "te", "kernel_wt", "param", "sza", "raa", "vza")}[self.kind]
        self.h2d = int(sum(np.asarray(sc[k]).size * 8 for k in used))
        self.d2h = int((7 if self.kind != "lambert" else (4 if self.S == 3 else 2)) * self.units_per_step * 8)
        self.dev_entry = self.ngeom == 1 and self.S == 3 and self.kind in ("lambert", "cx")
        if self.dev_entry:
            flat = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F")))
            self.d = {k: flat(sc[k]).to(dev) for k in used if k not in ("pe", "te")}
            self.d_out = [torch.empty(self.units_per_step, dtype=torch.float64, device=dev) for _ in range(5)]
        self.lib.ab200_vlidort_stats(None, None, None, None, None, None, 1)

    def _p(self, t):
        return self.C.cast(t.data_ptr(), self.C.POINTER(self.C.c_double))

    def step_gpu(self, stream):
        """device-resident inputs where a *_dev entry exists (vector Lambertian / Cox-Munk); otherwise the host entry (value == e2e)"""
        if not self.dev_entry:
            return self.step_e2e()
        C, d, p = self.C, self.d, self._p
        rc = C.c_int(0)
        if self.kind == "lambert":
            self.check(self.lib.ab200_vlidort_vector_lambert_dev(
                self.km, self.nch, self.nobs, self.N, int(self.plane_parallel), self.nmom, 6, p(d["ROT"]), p(d["depol"]), p(d["alpha"]),
                p(d["tau"]), p(d["ssa"]), p(d["pmom"]), p(d["he"]), p(d["albedo"]), p(d["sza"]), p(d["raa"]), p(d["vza"]), p(d["flux"]),
                C.c_double(1e15), *[p(o) for o in self.d_out[:4]], C.c_void_p(stream), C.byref(rc)), "vector_lambert_dev")
        else:
            self.check(self.lib.ab200_vlidort_vector_cx_dev(
                self.km, self.nch, self.nobs, self.N, int(self.plane_parallel), self.nmom, 6, p(d["ROT"]), p(d["depol"]), p(d["tau"]),
                p(d["ssa"]), p(d["pmom"]), p(d["he"]), p(d["u10m"]), p(d["v10m"]), p(d["mr"]), p(d["sza"]), p(d["raa"]), p(d["vza"]),
                C.c_double(1e15), *[p(o) for o in self.d_out], C.c_void_p(stream), C.byref(rc)), "vector_cx_dev")
        assert rc.value == 0

    def step_e2e(self):
        self.last = self.call_host(self.pinned)

    def step_pageable(self):
        self.last = self.call_host(self.scene)

    def stats(self, reset=False):
        C = self.C
        a, b, c = C.c_double(), C.c_double(), C.c_double(); d, e, f = C.c_long(), C.c_long(), C.c_long()
        fe, fb, nl, mo, fx = C.c_double(), C.c_double(), C.c_long(), C.c_double(), C.c_double()
        self.lib.ab200_vlidort_stats2(C.byref(fe), C.byref(fb), C.byref(nl))
        self.lib.ab200_vlidort_stats3(C.byref(mo), C.byref(fx))
        self.lib.ab200_vlidort_stats(C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e), C.byref(f), int(reset))
        return dict(ms_setup=a.value, ms_eigen=b.value, ms_bvp=c.value, n_eigen=d.value, n_bvp=e.value, npix=f.value,
                    flops_eigen=fe.value, flops_bvp=fb.value, n_launches=nl.value, ms_ops=mo.value, flops_exec=fx.value)

    # ---- CPU arms: the oracle restatement of the reference at the wrapper level, on the first `sample` observations of the same scene
    def setup_cpu(self, sample_cols):
        cores = os.cpu_count() or 1
        want = max(cores, -(-self.parity_pixels // (self.nch * self.ngeom)))
        self.sample_cols = max(1, min(self.nobs, sample_cols if 0 < sample_cols < 10 ** 8 else want))
        n = self.sample_cols
        if self.scene is not None:
            s = self.scene
            cut = {"ROT": 1, "alpha": 1, "tau": 2, "ssa": 2, "g": 2, "pmom": 2, "pe": 1, "he": 1, "te": 1, "albedo": 0, "sza": 0, "raa": 0, "vza": 0,
                   "flux": 1, "u10m": 0, "v10m": 0, "kernel_wt": 2, "param": 2}
            self.cpu_scene = {k: (np.asfortranarray(np.take(v, range(n), axis=cut[k])) if k in cut else v) for k, v in s.items()}
        else:
            self.cpu_scene = self.build_scene("oracle", cols=n)
        self.cpu_flat = self._flatten_views(self.cpu_scene)

    def _flatten_views(self, sc):
        """the oracle has no multi-geometry form: pixel-views as G * n independent observations, index g * n + j"""
        G = self.ngeom
        if G == 1:
            return sc
        n = sc["tau"].shape[2]
        rep = lambda a, ax: np.asfortranarray(np.concatenate([a] * G, axis=ax))
        f = dict(sc)
        f.update(ROT=rep(sc["ROT"], 1), alpha=rep(sc["alpha"], 1), tau=rep(sc["tau"], 2), ssa=rep(sc["ssa"], 2), g=rep(sc["g"], 2),
                 pmom=rep(sc["pmom"], 2), pe=rep(sc["pe"], 1), he=rep(sc["he"], 1), te=rep(sc["te"], 1), albedo=rep(sc["albedo"], 0),
                 flux=rep(sc["flux"], 1), sza=np.tile(sc["sza"], G), raa=sc["raa"].ravel(order="F"), vza=sc["vza"].ravel(order="F"),
                 u10m=np.tile(sc["u10m"], G), v10m=np.tile(sc["v10m"], G), kernel_wt=rep(sc["kernel_wt"], 2), param=rep(sc["param"], 2))
        return f

    def step_cpu(self, nthreads, tol=1e-12):
        from oracle import parity
        t0 = time.perf_counter()
        ref = parity.oracle_parallel(self.kind, self.S, self.N, self.plane_parallel, self.cpu_flat, nthreads, scalar_glue=(self.S == 1),
                                     asymtx_tol=tol)
        self.cpu_ref = ref
        return self.sample_cols * self.nch * self.ngeom, time.perf_counter() - t0

    def parity(self, nthreads):
        """the GPU answers of the e2e leg on the sampled observations against the oracle (protocol of oracle/parity.py)"""
        from oracle import parity
        n, G, nch = self.sample_cols, self.ngeom, self.nch
        L = self.last
        def cut(a):
            if a is None:
                return None
            a = np.asarray(a).reshape(self.nobs, nch, G, order="F")[:n]
            return np.asfortranarray(np.concatenate([a[:, :, g] for g in range(G)], axis=0))          # (G*n, nch), index g*n + j
        nf = np.asarray(L["nfourier"]).reshape(self.nobs, nch, G)[:n]
        got = dict(radiance=cut(L["radiance"]), reflectance=cut(L["reflectance"]), Q=cut(L["Q"]), U=cut(L["U"]),
                   nfourier=np.concatenate([nf[:, :, g] for g in range(G)], axis=0))
        return parity.parity_report(self.kind, self.S, self.N, self.plane_parallel, self.cpu_flat, got, nthreads, scalar_glue=(self.S == 1))

    def roofline(self, step_ms):
        st = self.stats()
        fp = os.path.join(ROOT, "profiles", "r01_fp64_peak.json")
        peak = json.load(open(fp))["dfma_tflops"] if os.path.exists(fp) else 37.0
        dom = "k_vl_eigen" if st["ms_eigen"] >= st["ms_bvp"] else "k_vl_layer_ops+k_vl_add"
        ms = st["ms_eigen"] if dom == "k_vl_eigen" else st["ms_bvp"]
        fl = st["flops_eigen"] if dom == "k_vl_eigen" else st["flops_bvp"]
        ach = fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        rt_ms = st["ms_eigen"] + st["ms_bvp"] + st["ms_setup"]
        tot = (st["flops_eigen"] + st["flops_bvp"]) / (rt_ms * 1e-3) / 1e12 if rt_ms > 0 else 0.0
        bvp_alg = st["flops_bvp"] / (st["ms_bvp"] * 1e-3) / 1e12 if st["ms_bvp"] > 0 else 0.0
        bvp_exec = st["flops_exec"] / (st["ms_bvp"] * 1e-3) / 1e12 if st["ms_bvp"] > 0 else 0.0
        traffic, tnote = None, None
        tp = os.path.join(ROOT, "profiles", "r02_rt_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp)).get(dom)
            if tj:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                tnote = tj.get("note")
        return {"bound": "fp64", "kernel": dom, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "traffic_note": tnote,
                "flops": "algorithmic count of SURVEY.md 8(d): eigen L (29 n^3 + 4 S^3 N^2 (2N - m)), boundary-value problem L ((14/3)(2n)^3 + 12 (2n)^2)",
                "peak_source": "measured DFMA microbenchmark on this pool (tools/fp64_peak.cu, profiles/r01_fp64_peak.json); "
                               "MEASURED_PEAKS.json holds no FP64 figure",
                "whole_rt_tflops": tot, "whole_rt_frac": tot / peak,
                "kernel_ms": {"setup": st["ms_setup"], "eigen": st["ms_eigen"], "layer_ops": st["ms_ops"], "add": st["ms_bvp"] - st["ms_ops"]},
                "bvp": {"algorithmic_tflops": bvp_alg, "frac": bvp_alg / peak, "executed_tflops": bvp_exec, "executed_frac": bvp_exec / peak,
                        "note": "the interaction-principle solver executes ~18 n^3 flops per layer where the band LU the algorithmic "
                                "count assumes needs 37 n^3; both figures are reported"},
                "fourier_terms_per_unit": st["n_bvp"] * self.ngeom / max(st["npix"], 1)}


class RtVectorCxWorkload(RtWorkload):
    """BASELINE.json configs[2] (headline): vector VLIDORT, 12 streams, 72 layers, Greek moments from the Mie tables (nMom 300), Cox-Munk
    ocean BRDF (f2py VECTOR_CX), leo_vlidort-style swath of synthetic pixels."""
    name = "vlidort_vector_n12_l72_coxmunk"
    metric = "VLIDORT pixels/sec (vector nstokes=3, 12 streams, 72 layers, nMom=300, FO + delta-M, Cox-Munk BRDF)"
    nobs, N, S, nmom, kind = 4096, 12, 3, 300, "cx"

    def pipeline_leg(self, torch, steps=2):
        """the same pixels through ab200_pipeline_vector_cx (getEdgeVars -> ROT_CALC -> getAOPvector -> VECTOR_CX in one call, the
        intermediates stay in HBM): host inputs are the model state.  Returns pixels/s and the bytes that cross PCIe."""
        from aeroapps_b200 import pipeline
        st, sc = self.state_n, self.scene
        a = (self.mie, [550.0], st["vnames"], st["qm"], st["rh"], st["airdens"], st["delp"], 1.0, self.N, self.plane_parallel,
             self.nmom, 6, sc["u10m"], sc["v10m"], sc["mr"], sc["sza"], sc["raa"], sc["vza"], 1e15, 0)
        out = pipeline.vector_cx_pipeline(*a)
        assert out[-1] == 0
        same = bool(np.array_equal(out[0], self.last["radiance"]) and np.array_equal(out[3], self.last["Q"]))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            pipeline.vector_cx_pipeline(*a)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        h2d = 4 * (st["qm"].size + st["rh"].size + st["airdens"].size + st["delp"].size) + 8 * 5 * self.nobs
        return {"value": self.units_per_step / dt, "unit": self.unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 5 * 8 * self.nobs,
                "identical_to_separate_calls": same,
                "entry": "ab200_pipeline_vector_cx (device-resident chain getEdgeVars -> ROT_CALC -> getAOPvector -> VECTOR_CX, per-rank wall clock)"}


class RtVectorWorkload(RtWorkload):
    name = "vlidort_vector_n12_l72_lambert"
    metric = "VLIDORT pixels/sec (vector nstokes=3, 12 streams, 72 layers, nMom=300, FO + delta-M, Lambertian surface)"
    nobs, N, S, nmom, kind = 4096, 12, 3, 300, "lambert"


class RtScalarWorkload(RtWorkload):
    """BASELINE.json configs[1]: scalar VLIDORT (nstokes = 1), 8 streams, 72 layers, Lambertian surface, nMom = 64, 1e5 pixels in steps of
    16384.  Only the reference-style host entry exists for the scalar form: `value` includes the PCIe copies and equals `e2e`."""
    name = "vlidort_scalar_n8_l72_lambert"
    metric = "VLIDORT pixels/sec (scalar nstokes=1, 8 streams, 72 layers, nMom=64, FO + delta-M, Lambertian surface)"
    nobs, N, S, nmom, kind = 16384, 8, 1, 64, "lambert"
    seed = 20260101 + 2
    parity_pixels = 1024


class RtGeoModisWorkload(RtWorkload):
    """BASELINE.json configs[3]: geo_vlidort TEMPO-like pixels, 4 channels (354/388/550/670 nm), vector 12 streams, MODIS RTLS (Ross-Li) BRDF
    (VLIDORT_Vector_LandMODIS); the 1e6-pixel disk is processed in steps of 1024 observations x 4 channels."""
    name = "geo_vlidort_modis_rtls_4ch_n12_l72"
    metric = "VLIDORT pixel-channels/sec (vector nstokes=3, 12 streams, 72 layers, nMom=300, 4 channels, MODIS RTLS BRDF)"
    unit = "pixel-channels/s"
    nobs, N, S, nmom, kind = 1024, 12, 3, 300, "modis"
    nch_nm = (354.0, 388.0, 550.0, 670.0)
    seed = 20260101 + 4


class RtPolarWorkload(RtWorkload):
    """BASELINE.json configs[4]: polar_vlidort multi-angle polarimeter, 60 view angles per column, nstokes 3, 16 streams, nMom 300, through the
    multi-geometry entry (one eigen / boundary-value solve per column and Fourier term for all 60 views); 256 columns per step."""
    name = "polar_vlidort_60views_n16_l72_lambert"
    metric = "VLIDORT pixel-views/sec (vector nstokes=3, 16 streams, 72 layers, nMom=300, 60 view angles per column, Lambertian)"
    unit = "pixel-views/s"
    nobs, N, S, nmom, kind, ngeom = 256, 16, 3, 300, "lambert", 60
    seed = 20260101 + 5
    parity_pixels = 240


WORKLOADS = {"aop_scalar": AopScalarWorkload, "rt_vector": RtVectorWorkload, "rt_vector_cx": RtVectorCxWorkload, "rt_scalar": RtScalarWorkload,
             "geo_modis": RtGeoModisWorkload, "polar16": RtPolarWorkload}
DEFAULT_WORKLOAD = "rt_vector_cx"
EXTRAS = ("rt_scalar", "geo_modis", "polar16")          # BASELINE configs[1], [3], [4]: attached to the default line


def time_aop_vector(wl, torch):
    """k_aop_vector (getAOPvector with the nMom x 6 phase-matrix moments) on device-resident inputs: voxels/s and the HBM bandwidth its
    algorithmic bytes (4 (nq+1) + 4 nch (3 + nMom nPol) per voxel) amount to"""
    import ctypes as C
    from aeroapps_b200._lib import lib, check
    dev = torch.device("cuda", torch.cuda.current_device())
    st, mie = wl.state_n, wl.mie
    nobs = st["qm"].shape[2]
    nq, nvox = len(st["vnames"]), wl.km * nobs
    d_qm = torch.from_numpy(np.ascontiguousarray(st["qm"].T)).to(dev); d_rh = torch.from_numpy(np.ascontiguousarray(st["rh"].T)).to(dev)
    outs = [torch.empty(nvox, dtype=torch.float32, device=dev) for _ in range(3)]
    d_pm = torch.empty(nvox * wl.nmom * 6, dtype=torch.float32, device=dev)
    ch = np.array([550.0], np.float32); vn = wl.mt.pad_names(st["vnames"])
    p = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_float))
    rc = C.c_int(0)
    stream = torch.cuda.current_stream().cuda_stream

    def run():
        check(lib().ab200_get_aop_vector_dev(mie.handle, wl.km, nobs, 1, nq, ch.ctypes.data_as(C.POINTER(C.c_float)),
                                             vn.ctypes.data_as(C.c_char_p), p(d_qm), p(d_rh), wl.nmom, 6, p(outs[0]), p(outs[1]),
                                             p(outs[2]), p(d_pm), C.c_void_p(stream), C.byref(rc)), "aop_vector_dev")
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    bpu = 4 * (nq + 1) + 4 * (3 + wl.nmom * 6)
    gbs = nvox * bpu / (ms * 1e-3) / 1e9
    pk = peaks()
    return {"kernel": "k_aop_vector", "voxels_per_s": nvox / (ms * 1e-3), "ms": ms, "bytes_per_unit": bpu, "achieved_gbs": gbs,
            "frac_of_hbm_peak": gbs / pk["hbm_gbs"], "voxels": nvox, "nMom": wl.nmom, "nPol": 6}


def time_angles(torch, n=4_000_000):
    """the geometry calculator of the geo driver (GeoAngles_.satangles) on n points: device-resident rate with its
    HBM roofline (92 B per point: five int32 + three float64 in, four float64 out), the rate through the host-pointer entry, the CPU
    restatement on one core, and the largest difference between the two on the sample"""
    import ctypes as C
    from aeroapps_b200._lib import lib, check
    from aeroapps_b200.angles import GeoAngles_
    rng = np.random.default_rng(20260107)
    ints = [rng.integers(lo, hi, n).astype(np.int32) for lo, hi in ((2005, 2030), (1, 13), (1, 29), (0, 24), (0, 60))]
    sec, lat, lon = rng.uniform(0, 60, n), rng.uniform(-70, 70, n), rng.uniform(-160, -40, n)
    dev = torch.device("cuda", torch.cuda.current_device())
    di = [torch.from_numpy(a).to(dev) for a in ints]; dd = [torch.from_numpy(a).to(dev) for a in (sec, lat, lon)]
    lss = torch.full((n,), -100.0, dtype=torch.float64, device=dev); las = torch.zeros(n, dtype=torch.float64, device=dev)
    outs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(4)]
    pi = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_int)); pd = lambda t: C.cast(t.data_ptr(), C.POINTER(C.c_double))
    stream = torch.cuda.current_stream().cuda_stream
    run = lambda: check(lib().ab200_satangles_dev(n, *[pi(t) for t in di], *[pd(t) for t in dd], pd(las), pd(lss), C.c_double(0.0),
                                                  C.c_double(35786.0), 0, *[pd(o) for o in outs], C.c_void_p(stream)), "satangles_dev")
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    bpu = 5 * 4 + 5 * 8 + 4 * 8                                  # the _dev form also reads per-point latss / lonss
    GeoAngles_.satangles(*[a[:1000] for a in ints], sec[:1000], lat[:1000], lon[:1000], -100.0)       # warm-up of the host-pointer entry
    th = 1e30
    for _ in range(2):                                            # best of two: one call is 50 ms, allocator noise is of that order
        t0 = time.perf_counter(); got = GeoAngles_.satangles(*ints, sec, lat, lon, -100.0); th = min(th, time.perf_counter() - t0)
    from oracle import angles_oracle as ao                       # CPU leg: the checker, timed as cpu_baseline
    m = 400_000
    t0 = time.perf_counter(); ref = ao.angles(0, *[a[:m] for a in ints], sec[:m], lat[:m], lon[:m], lonss0=-100.0); tc = time.perf_counter() - t0
    err = max(float(np.max(np.abs(np.asarray(g)[:m] - r))) for g, r in zip(got, ref))
    pk = peaks()
    gbs = n * bpu / (ms * 1e-3) / 1e9
    return {"metric": "GeoAngles satangles points/sec (view + solar azimuth / zenith per point)", "value": n / (ms * 1e-3), "unit": "points/s",
            "ms_per_step": ms, "points": n,
            "roofline": {"bound": "hbm", "kernel": "k_angles", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                         "bytes_per_unit": bpu, "traffic": None,
                         "note": "~30 FP64 transcendentals per point: the FP64 pipe, not HBM, is what binds this kernel"},
            "e2e": {"value": n / th, "unit": "points/s", "h2d_bytes_per_step": n * 44, "d2h_bytes_per_step": n * 32, "host_buffers": "pageable"},
            "cpu_baseline": {"value": m / tc, "unit": "points/s", "cores": 1, "kind": "port", "sample": f"{m} of {n} points, scalar C restatement"},
            "parity": {"n": m, "max_abs_deg": err, "bar_deg": 1e-6, "pass": bool(err <= 1e-6),
                       "note": "1e-9 degrees away from the formulation's singular points, conditioning-scaled near them: tests/test_angles.py"}}


def in_library_leg(world, torch):
    """strong scaling INSIDE the library (no torchrun ranks involved): one process, one call of the reference-style host entry
    VLIDORT_Vector_LandMODIS on a FIXED geo_vlidort workload (BASELINE configs[3] shape: 2048 observations x 4 channels), observations
    split into one slab per device by ab200_set_devices(world).  Rank 0 runs it while the other ranks wait."""
    from aeroapps_b200._lib import lib
    wl = RtGeoModisWorkload(nobs=2048)
    wl.setup_gpu(torch, torch.device("cuda", torch.cuda.current_device()))
    L = lib()
    used = L.ab200_set_devices(world)
    try:
        wl.step_e2e()                                          # warm-up: per-device workspaces, kernel attributes
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            wl.step_e2e()
        dt = (time.perf_counter() - t0) / 2
    finally:
        L.ab200_set_devices(1)
    return {"value": wl.units_per_step / dt, "unit": wl.unit, "devices": int(used), "scaling": "strong", "ms_per_call": 1e3 * dt,
            "units_per_call": wl.units_per_step, "config": wl.config(),
            "entry": "ab200_vlidort_vector_brdf_modis (host pointers, pinned) with ab200_set_devices(N): one host thread and one "
                     "observation slab per device, outputs copied by every device into its slab of the caller's arrays"}


def cpu_arm(wl, cores, passes=1, warmup=0):
    """the reference algorithm's CPU restatement (oracle/, built -O3 -march=native on this host for the timing) on `cores` threads over
    the workload's bounded sample; returns (units/s, sample description)"""
    native = False
    try:
        from oracle import vlidort_oracle as vo, mie_oracle as mo
        native = vo.use_native(True) if hasattr(wl, "kind") else mo.use_native(True)
    except Exception:
        native = False
    try:
        for _ in range(warmup):
            wl.step_cpu(cores)
        units = 0; secs = 0.0
        for _ in range(passes):
            u, dt = wl.step_cpu(cores)
            units += u; secs += dt
    finally:
        if native:
            vo.use_native(False); mo.use_native(False)
    what = wl.unit.split("/")[0]
    return units / secs, secs, (f"{units // max(passes, 1)} of {wl.units_per_step} {what} of the same step, {passes} pass(es), {cores} threads "
                                f"(oracle/: C/C++ restatement of the reference Fortran, "
                                f"{'-O3 -march=native build' if native else '-O2 build'}; thread pool over observation ranges)")


def run_reference(args, wl):
    """--impl reference: the reference algorithm's CPU implementation (oracle restatement; the Fortran itself cannot be built here or on
    the GPU box: no Fortran compiler, MAPL/ESMF un-vendored) on all host cores, on the SAME pixels as the GPU arm (same seeds; Stage A
    by the oracle, bit-identical to the product's)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    wl.setup_cpu(sample_cols=args.cpu_sample)
    v, secs, sample = cpu_arm(wl, cores, passes=args.steps, warmup=args.warmup)
    print(json.dumps({
        "impl": "reference", "metric": wl.metric, "value": v, "unit": wl.unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": wl.config(),
        "cpu_baseline": {"value": v, "unit": wl.unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_gpu(wl, args, torch, dev, dist, rank, world, local):
    """W warm-up + K timed steps of one workload (device-resident inputs where a device entry exists), then the e2e leg (pinned host
    buffers through the reference-style host entry) and, for Stage B, the same call with pageable host arrays; returns the JSON dict"""
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

    def timed(fn, nsteps, nwarm):
        for _ in range(nwarm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(nsteps):
            fn()
        e1.record()
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return wl.units_per_step * world * nsteps / (float(tt.item()) * 1e-3)

    # e2e leg: host (pinned) buffers through the host-pointer C-ABI call, copies inside the timed region
    e2e_steps = max(3, min(args.steps, 10))
    io = None
    if isinstance(wl, RtWorkload):
        # the library counts what its host-pointer entries really move (of pmom only the elements and moments the solve reads are uploaded)
        import ctypes as C
        for _ in range(2):
            wl.step_e2e()
        wl.lib.ab200_vlidort_io_bytes(None, None, 1)
        e2e_value = timed(wl.step_e2e, e2e_steps, 0)
        h, d = C.c_double(), C.c_double()
        wl.lib.ab200_vlidort_io_bytes(C.byref(h), C.byref(d), 1)
        io = (int(h.value / e2e_steps), int(d.value / e2e_steps))
    else:
        e2e_value = timed(wl.step_e2e, e2e_steps, 2)
    e2e = {"value": e2e_value, "unit": wl.unit, "h2d_bytes_per_step": io[0] if io else wl.h2d, "d2h_bytes_per_step": io[1] if io else wl.d2h,
           "host_buffers": "pinned"}
    if io:
        e2e["host_array_bytes_per_step"] = wl.h2d              # size of the arrays the caller passes (the reference's argument list)
    if hasattr(wl, "step_pageable"):
        e2e["pageable_host_value"] = timed(wl.step_pageable, 2, 1)     # what a Fortran / numpy caller with ordinary arrays gets
    if rank != 0:
        return None
    return {
        "metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": wl.config(), "roofline": roof,
        "e2e": e2e, "gpu_launches": int(launches), "clocks": cs.summary(),
    }


def cpu_and_parity(wl, out, args):
    """rank 0, one GPU: time the CPU oracle on the bounded sample (cpu_baseline) and compare the GPU answers of the e2e leg with it"""
    cores = os.cpu_count() or 1
    wl.setup_cpu(sample_cols=args.cpu_sample)
    v, secs, sample = cpu_arm(wl, cores)
    out["cpu_baseline"] = {"value": v, "unit": wl.unit, "cores": cores, "kind": "port", "sample": sample}
    if hasattr(wl, "parity"):
        out["parity"] = wl.parity(cores)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--nobs", type=int, default=0, help="override columns/pixels per GPU per step (debug)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="observations in the bounded CPU sample (0 = workload default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    ap.add_argument("--no-aop", action="store_true", help="skip the extra Mie-AOP measurement attached to the RT line")
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configurations attached to the default line")
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
    if dist and getattr(wl, "last", None) is not None:
        # the path's only exchange: the final gather of the per-rank slabs (outside the timed steps)
        from aeroapps_b200 import swath
        loc = np.ascontiguousarray(np.asarray(wl.last["radiance"]).reshape(-1))
        full = swath.gather_swath(loc, loc.shape[0] * world, axis=0, dist=dist, device=dev)
        if rank == 0:
            out["config"]["gathered_outputs"] = int(full.shape[0])
            out["config"]["parallelism"] = f"{world} ranks x contiguous slabs, no data-path collective, final all_gather over NCCL"
    if hasattr(wl, "pipeline_leg"):
        pl = wl.pipeline_leg(torch)
        if rank == 0:
            out["pipeline"] = pl
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_and_parity(wl, out, args)
    is_default = args.workload == DEFAULT_WORKLOAD and not args.nobs
    if args.workload.startswith("rt_vector") and not args.no_aop:
        # the second half of the BASELINE metric: Mie AOP voxels/s (BASELINE configs[0], c90 scalar AOPs)
        import copy
        a2 = copy.copy(args); a2.steps = 20; a2.warmup = 5
        awl = AopScalarWorkload()
        aop = run_gpu(awl, a2, torch, dev, dist, rank, world, local)
        if rank == 0:
            out["aop"] = {k: aop[k] for k in ("metric", "value", "unit", "ms_per_step", "roofline", "e2e", "config")}
            out["aop"]["vector_e2e_voxels_per_s_nmom300"] = getattr(wl, "aop_e2e_voxels_per_s", None)
            if hasattr(wl, "mie"):
                out["aop"]["vector"] = time_aop_vector(wl, torch)
            if world == 1 and not args.no_cpu:
                a2.cpu_sample = 48600
                cpu_and_parity(awl, out["aop"], a2)
    if is_default and not args.no_extras:
        # the other BASELINE configurations, each with its own value / e2e / roofline / cpu_baseline / parity (short legs)
        import copy
        del wl                      # frees the headline scene (pinned + device copies) before the next one is built
        import gc; gc.collect(); torch.cuda.empty_cache()
        for name in EXTRAS:
            a3 = copy.copy(args); a3.steps = 2; a3.warmup = 3; a3.cpu_sample = 10 ** 9
            xwl = WORKLOADS[name]()
            x = run_gpu(xwl, a3, torch, dev, dist, rank, world, local)
            if rank == 0:
                if world == 1 and not args.no_cpu:
                    cpu_and_parity(xwl, x, a3)
                out.setdefault("extras", {})[name] = x
            del xwl
            gc.collect(); torch.cuda.empty_cache()
        if rank == 0 and world == 1:
            out["extras"]["geo_angles"] = time_angles(torch)
        # in-library multi-GPU: rank 0 alone drives all `world` devices through ONE library call.  The other ranks must leave their
        # GPUs idle meanwhile, so they wait on the host (rendezvous-store key), not inside an NCCL barrier kernel.
        if dist:
            dist.barrier()
            torch.cuda.synchronize()
            store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            out["in_library_multi_gpu"] = in_library_leg(world, torch)
            if dist:
                store.set("ab200_in_library_done", "1")
        elif dist:
            store.wait(["ab200_in_library_done"])
        if dist:
            dist.barrier()
    if rank == 0:
        print(json.dumps(out))
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
