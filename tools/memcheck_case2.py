import sys; sys.path.insert(0, '.')
import numpy as np
from aeroapps_b200 import VLIDORT_POLAR_ as VP
from oracle import vlidort_glue as vg
S, N, nobs = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sc = vg.synthetic_scene(nobs, nmom=64 if S == 1 else 60, seed=48)
if S == 1:
    I, R, rc = VP.scalar_lambert(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["g"], sc["pmom"], sc["pe"], sc["he"], sc["te"], sc["albedo"], sc["sza"], sc["raa"], sc["vza"], 1e15, 0)
else:
    I, R, BR, Q, U, _, _, rc = VP.vector_cx(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"], sc["te"], sc["u10m"], sc["v10m"], sc["mr"], sc["sza"], sc["raa"], sc["vza"], 1e15, 0)
print("rc", rc, I.ravel()[:6])
nf, st = VP.last_diag()
print("status", st.ravel()[:8], "nf", nf.ravel()[:8])
