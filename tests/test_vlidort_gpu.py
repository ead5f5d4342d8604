"""GPU parity tests for Stage B: the CUDA path, called through the C ABI (ctypes mirror VLIDORT_POLAR_), against
the CPU oracle (pinned to VLIDORT's saved results by tests/test_vlidort_oracle_golden.py) on the same seeded
synthetic pixels.  Bar (north_star): Stokes I/Q/U and reflectance <= 1e-8 relative (Q,U relative to I), identical
Fourier counts, identical MISSING / status behaviour."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1.0e-8


def rel(a, b, scale):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(scale), 1e-300)))


@pytest.fixture(scope="module")
def mods(built):
    from aeroapps_b200 import VLIDORT_POLAR_ as VP
    from oracle import vlidort_glue as vg
    return VP, vg


def test_rot_calc(mods):
    VP, vg = mods
    sc = vg.synthetic_scene(7, nch_nm=(354.0, 550.0, 870.0), nmom=4)
    ROT, depol, rc = VP.rot_calc(sc["channels"], sc["pe"], sc["he"], sc["te"], 1e15, 0)
    assert rc == 0
    # double precision with FMA contraction on the device: a few ulp
    assert rel(ROT, sc["ROT"], sc["ROT"]) < 1e-12 and rel(depol, sc["depol"], sc["depol"]) < 1e-12


@pytest.mark.parametrize("N,planpar,nobs,nch", [(8, False, 33, (550.0,)), (8, True, 9, (550.0, 870.0)), (4, False, 5, (550.0,))])
def test_scalar_lambert(mods, N, planpar, nobs, nch):
    VP, vg = mods
    sc = vg.synthetic_scene(nobs, nch_nm=nch, nmom=64, seed=40 + N)
    sc["albedo"][0, 0] = 0.0; sc["vza"][1] = 0.0
    ref = vg.vlidort_run(1, N, planpar, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["he"], sc["sza"], sc["raa"],
                         sc["vza"], albedo=sc["albedo"], scalar_glue=True)
    I, R, rc = VP.scalar_lambert(sc["channels"], N, planpar, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["g"], sc["pmom"],
                                 sc["pe"], sc["he"], sc["te"], sc["albedo"], sc["sza"], sc["raa"], sc["vza"], 1e15, 0)
    assert rc == 0
    nf, st = VP.last_diag()
    assert (st == 0).all()
    assert np.array_equal(nf.reshape(nobs, len(nch)), ref["nfourier"])
    assert rel(I, ref["radiance"], ref["radiance"]) < TOL
    assert rel(R, ref["reflectance"], ref["reflectance"]) < TOL


@pytest.mark.parametrize("N,planpar,nobs", [(6, False, 12), (12, True, 6), (4, False, 3)])
def test_vector_lambert(mods, N, planpar, nobs):
    VP, vg = mods
    sc = vg.synthetic_scene(nobs, nmom=80, seed=60 + N)
    rng = np.random.default_rng(1)
    alpha = np.asfortranarray(rng.uniform(0, 0.01, sc["ROT"].shape))
    flux = np.asfortranarray(rng.uniform(0.5, 2.0, (1, nobs)))
    ref = vg.vlidort_run(3, N, planpar, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["he"], sc["sza"], sc["raa"],
                         sc["vza"], albedo=sc["albedo"], alpha=alpha, flux_factor=flux)
    I, R, Q, U, rc = VP.vector_lambert(sc["channels"], N, planpar, sc["ROT"], sc["depol"], alpha, sc["tau"], sc["ssa"], sc["pmom"],
                                       sc["pe"], sc["he"], sc["te"], sc["albedo"], sc["sza"], sc["raa"], sc["vza"], flux, 1e15, 0)
    assert rc == 0
    nf, st = VP.last_diag()
    assert (st == 0).all() and np.array_equal(nf.reshape(nobs, 1), ref["nfourier"])
    Iref = ref["radiance"]
    assert rel(I, Iref, Iref) < TOL and rel(R, ref["reflectance"], ref["reflectance"]) < TOL
    assert rel(Q, ref["Q"], Iref) < TOL and rel(U, ref["U"], Iref) < TOL


def test_missing_and_bad_inputs(mods):
    VP, vg = mods
    M = 1e15
    sc = vg.synthetic_scene(6, nch_nm=(550.0, 670.0), nmom=16, seed=9)
    sc["sza"][1] = M                         # missing angle -> whole row MISSING
    sc["albedo"][2, 1] = M                   # missing albedo -> that (obs, channel) MISSING
    sc["sza"][3] = 93.0                      # VLIDORT input check fails -> MISSING, rc stays 0
    sc["sza"][4] = 0.0                       # FOCORR_NADIR state quirk from obs 4 on
    ref = vg.vlidort_run(1, 4, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["he"], sc["sza"], sc["raa"],
                         sc["vza"], albedo=sc["albedo"], scalar_glue=True)
    I, R, rc = VP.scalar_lambert(sc["channels"], 4, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["g"], sc["pmom"],
                                 sc["pe"], sc["he"], sc["te"], sc["albedo"], sc["sza"], sc["raa"], sc["vza"], M, 0)
    assert rc == 0
    miss_ref = ref["radiance"] == M
    assert np.array_equal(I == M, miss_ref) and np.array_equal(R == M, miss_ref)
    assert miss_ref[1].all() and miss_ref[2, 1] and miss_ref[3].all() and miss_ref.sum() == 5
    ok = ~miss_ref
    assert rel(I[ok], ref["radiance"][ok], ref["radiance"][ok]) < TOL
    # dimension checks of VLIDORT_Init
    _, _, rc = VP.scalar_lambert(sc["channels"], 26, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["g"], sc["pmom"],
                                 sc["pe"], sc["he"], sc["te"], sc["albedo"], sc["sza"], sc["raa"], sc["vza"], M, 0)
    assert rc == 2


def test_size_independent_properties(mods):
    """larger batch: (1) results do not depend on batch composition (a pixel alone == the pixel inside a batch),
    (2) flux linearity: I scales exactly with flux_factor while reflectance is invariant."""
    VP, vg = mods
    nobs = 257
    sc = vg.synthetic_scene(nobs, nmom=32, seed=77)
    alpha = np.zeros_like(sc["ROT"]); f1 = np.ones((1, nobs), order="F"); f2 = np.full((1, nobs), 2.0, order="F")
    args = lambda ff: (sc["channels"], 4, False, sc["ROT"], sc["depol"], alpha, sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"],
                       sc["te"], sc["albedo"], sc["sza"], sc["raa"], sc["vza"], ff, 1e15, 0)
    I1, R1, Q1, U1, _ = VP.vector_lambert(*args(f1))
    I2, R2, Q2, U2, _ = VP.vector_lambert(*args(f2))
    assert rel(I2, 2.0 * I1, I1) < 1e-13 and rel(R2, R1, R1) < 1e-13 and rel(Q2, 2.0 * Q1, I1) < 1e-13
    j = 100
    Ia, Ra, Qa, Ua, _ = VP.vector_lambert(sc["channels"], 4, False, sc["ROT"][:, j:j + 1], sc["depol"], alpha[:, j:j + 1],
                                          sc["tau"][:, :, j:j + 1], sc["ssa"][:, :, j:j + 1], sc["pmom"][:, :, j:j + 1],
                                          sc["pe"][:, j:j + 1], sc["he"][:, j:j + 1], sc["te"][:, j:j + 1], sc["albedo"][j:j + 1],
                                          sc["sza"][j:j + 1], sc["raa"][j:j + 1], sc["vza"][j:j + 1], f1[:, j:j + 1], 1e15, 0)
    assert Ia[0, 0] == I1[j, 0] and Qa[0, 0] == Q1[j, 0] and Ua[0, 0] == U1[j, 0]
