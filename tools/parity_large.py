"""One-off: the parity protocol of oracle/parity.py on a LARGE sample of the headline shape (vector, 12 streams, 72 layers, nMom 300, Cox-Munk),
counting Fourier-count flips and pixels above 1e-8.  usage: python tools/parity_large.py [nobs] > profiles/r02_parity/headline_cx_large.json"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
g.build()
from aeroapps_b200 import VLIDORT_POLAR_ as VP
from oracle import vlidort_glue as vg, parity, vlidort_oracle as vo
if hasattr(vo, "use_native"):
    vo.use_native()
nobs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
sc = vg.synthetic_scene(nobs, nmom=300, seed=20260201)
t0 = time.time()
I, R, BR, Q, U, BQ, BU, rc = VP.vector_cx(sc["channels"], 12, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"], sc["te"],
                                          sc["u10m"], sc["v10m"], sc["mr"], sc["sza"], sc["raa"], sc["vza"], 1e15, 0)
nf, st = VP.last_diag()
t1 = time.time()
rep = parity.parity_report("cx", 3, 12, False, sc, dict(radiance=I, reflectance=R, Q=Q, U=U, nfourier=nf.reshape(nobs, 1)))
rep.update(nobs=nobs, gpu_seconds=t1 - t0, oracle_seconds=time.time() - t1, fourier_histogram={int(k): int(v) for k, v in zip(*np.unique(nf, return_counts=True))})
print(json.dumps(rep, indent=1))
