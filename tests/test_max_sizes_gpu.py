"""Maximum sizes (VD/vlidort_pars.f90: MAXSTREAMS 25, MAXLAYERS 72).  The library covers nstokes * nstreams <= 64 with 72 layers: the scalar
forms up to the reference's 25 streams, the vector forms (nstokes 3) up to 21 streams -- 63 x 63 layer problems through the generic
one-warp K2 and the interaction-principle K3.  Both ends are run against the oracle (bar and protocol of oracle/parity.py).  Vector calls
with 22 - 25 streams, which the reference build accepts, are refused loudly (negative library error, nothing computed): the boundary-value
working set of one pixel no longer fits the 227 KB of shared memory of an SM in either solver."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
M = 1e15


@pytest.mark.parametrize("S,N", [(1, 25), (3, 21)])
def test_maximum_stream_counts(built, S, N):
    from aeroapps_b200 import VLIDORT_POLAR_ as VP
    from oracle import vlidort_glue as vg, parity
    nobs = 2
    sc = vg.synthetic_scene(nobs, nmom=80, seed=900 + N)
    assert sc["tau"].shape[0] == 72
    if S == 1:
        I, R, rc = VP.scalar_lambert(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["g"], sc["pmom"], sc["pe"], sc["he"],
                                     sc["te"], sc["albedo"], sc["sza"], sc["raa"], sc["vza"], M, 0)
        got = dict(radiance=I, reflectance=R, Q=None, U=None)
        kind = "lambert"
    else:
        I, R, BR, Q, U, _, _, rc = VP.vector_cx(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"],
                                                sc["te"], sc["u10m"], sc["v10m"], sc["mr"], sc["sza"], sc["raa"], sc["vza"], M, 0)
        got = dict(radiance=I, reflectance=R, Q=Q, U=U)
        kind = "cx"
    assert rc == 0
    nf, st = VP.last_diag()
    assert (st == 0).all(), st
    got["nfourier"] = nf.reshape(nobs, 1)
    rep = parity.parity_report(kind, S, N, False, sc, got, scalar_glue=(S == 1))
    print("parity max sizes", S, N, rep)
    assert rep["n"] == nobs and rep["pass"], str(rep)


@pytest.mark.parametrize("N", [22, 25])
def test_vector_beyond_21_streams_is_refused_loudly(built, N):
    from aeroapps_b200 import VLIDORT_POLAR_ as VP
    from oracle import vlidort_glue as vg
    sc = vg.synthetic_scene(2, nmom=80, seed=930)
    with pytest.raises(RuntimeError, match="too large"):
        VP.vector_cx(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"], sc["te"], sc["u10m"],
                     sc["v10m"], sc["mr"], sc["sza"], sc["raa"], sc["vza"], M, 0)
