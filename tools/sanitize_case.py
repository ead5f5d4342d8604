"""tiny Stage A + Stage B call for compute-sanitizer (memcheck / racecheck): 2 pixels, nstokes 3, 12 streams (two-warp eigen kernel,
default K3), Cox-Munk, plus the scalar and vector AOP kernels.  usage: compute-sanitizer --tool racecheck python tools/sanitize_case.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aeroapps_b200 import MieObs_, VLIDORT_POLAR_ as VP, mietables as mt
F = lambda a: np.asfortranarray(a, dtype=np.float64)
nobs, N, nmom = 2, int(os.environ.get("SAN_N", "12")), 40
tables = mt.synthetic_tables(nmom=nmom)
st = mt.synthetic_state(nobs, 72, seed=5)
mie = MieObs_.ChemMie([550e-9], nmom, tables)
tau, ssa, g, pmom, rc = mie.getaopvector([550.0], st["vnames"], 0, st["qm"], st["rh"], nmom, 6)
t2, s2, g2, rc2 = mie.getaopscalar([550.0], st["vnames"], 0, st["qm"], st["rh"])
pe, ze, te = MieObs_.getedgevars(st["airdens"], st["delp"], 1.0)
pm = pmom.astype(np.float64); pm[..., 1] *= -1.0; pm[..., 3] *= -1.0
ch = np.array([550.0])
ROT, depol, rc = VP.rot_calc(ch, F(pe), F(ze), F(te), 1e15, 0)
out = VP.vector_cx(ch, N, False, F(ROT), depol, F(tau), F(ssa), F(pm), F(pe), F(ze), F(te), np.array([3.0, -2.0]), np.array([1.0, 4.0]),
                   np.array([1.334]), np.array([30.0, 55.0]), np.array([20.0, 120.0]), np.array([10.0, 40.0]), 1e15, 0)
print("rc", out[-1], "I", out[0].ravel(), "nf", VP.last_diag()[0])
