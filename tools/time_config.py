"""wall-clock rate of one reference-style host call for BASELINE configurations that are not bench workloads:
   python tools/time_config.py polar16 [nobs]   (configs[4] shape: nstokes 3, 16 streams, Lambertian)
   python tools/time_config.py modis4 [nobs]    (configs[3] shape: nstokes 3, 12 streams, 4 channels, MODIS RTLS)"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aeroapps_b200 import MieObs_, VLIDORT_POLAR_ as VP, mietables as mt
F = lambda a: np.asfortranarray(a, dtype=np.float64)
kind = sys.argv[1]; nobs = int(sys.argv[2]) if len(sys.argv) > 2 else 512
nmom = 300
chn = [550.0] if kind == "polar16" else [354.0, 388.0, 550.0, 670.0]
tables = mt.synthetic_tables(nmom=nmom)
st = mt.synthetic_state(nobs, 72, seed=11)
mie = MieObs_.ChemMie([c * 1e-9 for c in chn], nmom, tables)
tau, ssa, g, pmom, rc = mie.getaopvector(chn, st["vnames"], 0, st["qm"], st["rh"], nmom, 6); assert rc == 0
pe, ze, te = MieObs_.getedgevars(st["airdens"], st["delp"], 1.0)
pm = pmom.astype(np.float64); pm[..., 1] *= -1.0; pm[..., 3] *= -1.0
ch = np.array(chn)
ROT, depol, rc = VP.rot_calc(ch, F(pe), F(ze), F(te), 1e15, 0)
rng = np.random.default_rng(3)
sza, vza, raa = rng.uniform(0, 75, nobs), rng.uniform(0, 65, nobs), rng.uniform(0, 180, nobs)
nch = len(chn)
def call():
    if kind == "polar16":
        alb = F(rng.uniform(0.02, 0.4, (nobs, nch)))
        return VP.vector_lambert(ch, 16, False, F(ROT), depol, np.zeros_like(F(ROT)), F(tau), F(ssa), F(pm), F(pe), F(ze), F(te), alb, sza, raa, vza,
                                 np.ones((nch, nobs), order="F"), 1e15, 0)
    kw = np.zeros((3, nch, nobs), order="F"); kw[0] = rng.uniform(.02, .4, (nch, nobs)); kw[1] = rng.uniform(0, .08, (nch, nobs)); kw[2] = rng.uniform(0, .2, (nch, nobs))
    par = np.zeros((2, nch, nobs), order="F"); par[0] = 2.0; par[1] = 1.0
    return VP.vector_brdf_modis(ch, 12, False, F(ROT), depol, F(tau), F(ssa), F(pm), F(pe), F(ze), F(te), kw, par, sza, raa, vza, 1e15, 0)
out = call(); assert out[-1] == 0
t0 = time.perf_counter(); out = call(); dt = time.perf_counter() - t0
nf, stt = VP.last_diag()
print(json.dumps({"config": kind, "nobs": nobs, "nch": nch, "pixel_channels_per_s": nobs * nch / dt, "seconds": dt,
                  "fourier_terms_per_pixel": float(np.mean(nf)), "not_run_minus500_or_missing": int(np.sum(stt != 0)), "stats": VP.stats(True) if hasattr(VP, "stats") else None}))
