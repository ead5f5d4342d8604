python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python - <<'PY' 2>&1 | tail -20
import time, numpy as np
from aeroapps_b200 import VLIDORT_POLAR_ as VP
from oracle import vlidort_glue as vg
for (S,N,nobs,nmom) in [(1,8,2048,64),(3,12,512,300)]:
    sc = vg.synthetic_scene(nobs, nmom=nmom, seed=1)
    alpha = np.zeros_like(sc["ROT"]); ff = np.ones((1,nobs),order="F")
    for rep in range(2):
        VP.stats(reset=True); t0=time.time()
        if S==1: out = VP.scalar_lambert(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["g"], sc["pmom"], sc["pe"], sc["he"], sc["te"], sc["albedo"], sc["sza"], sc["raa"], sc["vza"], 1e15, 0)
        else: out = VP.vector_lambert(sc["channels"], N, False, sc["ROT"], sc["depol"], alpha, sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"], sc["te"], sc["albedo"], sc["sza"], sc["raa"], sc["vza"], ff, 1e15, 0)
        dt=time.time()-t0
        nf, st = VP.last_diag()
        print(S,N,nobs,"wall s",round(dt,3),"pix/s",round(nobs/dt,1),VP.stats(),"nf hist",np.bincount(nf).tolist())
PY
