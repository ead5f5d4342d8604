"""GPU test of the chained column pipeline (ab200_pipeline_vector_cx): one call must give bit for bit
what the reference's call sequence getEdgeVars -> ROT_CALC -> getAOPvector -> VECTOR_CX gives through the separate
drop-in entry points (which the other GPU tests hold against the oracle)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nobs,N,nmom,wl", [(6, 6, 40, 469.3), (3, 12, 300, 550.0)])   # 469.3 nm is not a float32 number
def test_pipeline_matches_separate_calls(built, nobs, N, nmom, wl):
    from aeroapps_b200 import MieObs_, VLIDORT_POLAR_ as VP, mietables as mt, pipeline
    F = lambda a: np.asfortranarray(a, dtype=np.float64)
    tables = mt.synthetic_tables(nmom=nmom)
    st = mt.synthetic_state(nobs, 72, seed=31 + N)
    mie = MieObs_.ChemMie([wl * 1e-9], nmom, tables)
    ch = np.array([wl])
    rng = np.random.default_rng(9)
    sza, vza, raa = rng.uniform(5, 70, nobs), rng.uniform(0, 60, nobs), rng.uniform(0, 180, nobs)
    u10, v10, mr = rng.normal(0, 5, nobs), rng.normal(0, 5, nobs), np.array([1.334])
    u10[1] = 1e15                                            # a MISSING wind: that observation comes back as MISSING
    # the reference driver's sequence
    pe, ze, te = MieObs_.getedgevars(st["airdens"], st["delp"], 1.0)
    ROT, depol, rc = VP.rot_calc(ch, F(pe), F(ze), F(te), 1e15, 0)
    tau, ssa, g, pmom, rc = mie.getaopvector([wl], st["vnames"], 0, st["qm"], st["rh"], nmom, 6)
    assert rc == 0
    pm = pmom.astype(np.float64); pm[..., 1] *= -1.0; pm[..., 3] *= -1.0
    ref = VP.vector_cx(ch, N, False, F(ROT), depol, F(tau), F(ssa), F(pm), F(pe), F(ze), F(te), u10, v10, mr, sza, raa, vza, 1e15, 0)
    nf_ref, _ = VP.last_diag()
    out = pipeline.vector_cx_pipeline(mie, [wl], st["vnames"], st["qm"], st["rh"], st["airdens"], st["delp"], 1.0, N, False, nmom, 6,
                                      u10, v10, mr, sza, raa, vza, 1e15, 0)
    nf, _ = VP.last_diag()
    assert out[-1] == 0 and ref[-1] == 0 and np.array_equal(nf, nf_ref)
    for a, b, name in zip(out[:-1], ref[:-1], ("radiance", "reflectance", "BR", "Q", "U", "BR_Q", "BR_U")):
        assert np.array_equal(a, b), name
    assert out[0][1, 0] == 1e15 and out[0][0, 0] > 0
