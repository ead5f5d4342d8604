"""Cloud variants: the entry points whose reference wrapper calls VLIDORT_Run_Vector_Cloud
(V283/VLIDORT_ScatMod.F90:1865-2220) -- VLIDORT_Vector_GissCx_Cloud (GV/VLIDORT_BRDF_CX.F90:2276) and VLIDORT_Vector_LandMODIS_Cloud
(GV/VLIDORT_BRDF_MODIS.F90:825; f2py VECTOR_BRDF_MODIS_CLOUD) -- through the C ABI against the oracle with the same cloud glue
(oracle/vlidort_oracle.cpp, orc_set_cloud).  Bar and protocol as for the clear-sky entries (oracle/parity.py): 1e-8, identical Fourier
counts.  Also: zero cloud optical depth must reproduce the clear-sky entry (to 1e-12) when depol comes from ROT_CALC, and the depol
argument must not matter (the cloud runner recomputes it from the wavelength)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
M = 1e15


def _cloud_scene(vg, nobs, nch_nm, nmom, seed):
    sc = vg.synthetic_scene(nobs, nch_nm=nch_nm, nmom=nmom, seed=seed)
    rng = np.random.default_rng(seed + 1)
    km, nch, _ = sc["tau"].shape
    k = np.arange(km)[:, None, None]
    # a liquid cloud deck around layer 60 and a thin cirrus around layer 30, in two thirds of the columns
    on = (rng.uniform(size=(1, 1, nobs)) < 0.67)
    sc["tauL"] = np.asfortranarray(on * rng.uniform(0.2, 6.0, (1, nch, nobs)) * np.exp(-0.5 * ((k - 60) / 2.0) ** 2))
    sc["tauI"] = np.asfortranarray(on * rng.uniform(0.02, 0.8, (1, nch, nobs)) * np.exp(-0.5 * ((k - 30) / 3.0) ** 2))
    sc["ssaL"] = np.asfortranarray(np.full((km, nch, nobs), 0.99999)); sc["ssaI"] = np.asfortranarray(np.full((km, nch, nobs), 0.9997))
    l = np.arange(nmom)
    hg = lambda g: (2 * l + 1) * g ** l                                    # Henyey-Greenstein moments, P11 only is used
    pm = np.zeros((km, nch, nobs, nmom, 6))
    pL = pm.copy(); pL[..., 0] = hg(0.85); pL[..., 1:] = rng.uniform(-1, 1, (1, 1, 1, nmom, 5))     # the other elements must be ignored
    pI = pm.copy(); pI[..., 0] = hg(0.75); pI[..., 1:] = rng.uniform(-1, 1, (1, 1, 1, nmom, 5))
    sc["pmomL"] = np.asfortranarray(pL); sc["pmomI"] = np.asfortranarray(pI)
    sc["alpha"] = np.asfortranarray(rng.uniform(0.0, 0.01, sc["ROT"].shape)); sc["flux"] = np.asfortranarray(rng.uniform(0.9, 1.1, (nch, nobs)))
    return sc


@pytest.fixture(scope="module")
def mods(built):
    from aeroapps_b200 import VLIDORT_POLAR_ as VP
    from oracle import vlidort_glue as vg, parity
    return VP, vg, parity


def test_vector_gisscx_cloud(mods):
    VP, vg, parity = mods
    nobs, N = 12, 8
    sc = _cloud_scene(vg, nobs, (550.0,), 80, 501)
    cl = [sc[k] for k in ("tauI", "ssaI", "pmomI", "tauL", "ssaL", "pmomL")]
    args = lambda dp: (sc["channels"], N, False, sc["ROT"], dp, sc["tau"], sc["ssa"], sc["pmom"], *cl, sc["pe"], sc["he"], sc["te"], sc["u10m"],
                       sc["v10m"], sc["mr"], sc["sza"], sc["raa"], sc["vza"], M, 0)
    I, R, BR, Q, U, BQ, BU, rc = VP.vector_gisscx_cloud(*args(sc["depol"]))
    assert rc == 0
    nf, st = VP.last_diag()
    assert (st == 0).all()
    got = dict(radiance=I, reflectance=R, Q=Q, U=U, nfourier=nf.reshape(nobs, 1))
    flat = dict(sc); del flat["alpha"], flat["flux"]
    # liquid cloud layers of optical depth up to 6 with single scattering albedo 0.99999 sit on the runner's omega clamp: the smallest
    # eigenvalue of those layers is ~1e-2, the problem loses ~5 digits, and two exact-arithmetic-equivalent FP64 solvers differ by ~1e-9
    # (measured 1.1e-9 against the converged oracle, 7.9e-8 against the reference-tolerance one); converged bar 5e-9 here
    rep = parity.parity_report("gisscx", 3, N, False, flat, got, tol_conv=5e-9)
    print("parity gisscx_cloud", rep)
    assert rep["n"] == nobs and rep["pass"], str(rep)
    # the clouds matter (this is not the clear-sky answer) ...
    I0 = VP.vector_gisscx(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"], sc["te"],
                          sc["u10m"], sc["v10m"], sc["mr"], sc["sza"], sc["raa"], sc["vza"], M, 0)[0]
    assert np.max(np.abs(I - I0) / I0) > 0.05
    # ... the depol argument does not (recomputed from the wavelength) ...
    I2 = VP.vector_gisscx_cloud(*args(sc["depol"] * 0.5))[0]
    assert np.array_equal(I, I2)
    # ... and with no cloud optical depth the clear-sky entry comes back bit for bit (depol from ROT_CALC == coef_depol)
    z = np.zeros_like(sc["tau"])
    a0 = (sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], z, sc["ssaI"], sc["pmomI"], z, sc["ssaL"], sc["pmomL"],
          sc["pe"], sc["he"], sc["te"], sc["u10m"], sc["v10m"], sc["mr"], sc["sza"], sc["raa"], sc["vza"], M, 0)
    I3 = VP.vector_gisscx_cloud(*a0)[0]
    assert np.max(np.abs(I3 - I0) / I0) < 1e-12          # (the depolarisation ratio is recomputed on the host: last-bit differences)


def test_vector_brdf_modis_cloud(mods):
    VP, vg, parity = mods
    nobs, N = 10, 6
    sc = _cloud_scene(vg, nobs, (470.0, 660.0), 60, 502)
    sc["vza"][3] = M                                                            # a MISSING observation
    cl = [sc[k] for k in ("tauI", "ssaI", "pmomI", "tauL", "ssaL", "pmomL")]
    I, R, BR, Q, U, BQ, BU, rc = VP.vector_brdf_modis_cloud(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["alpha"], sc["tau"], sc["ssa"],
                                                            sc["pmom"], *cl, sc["pe"], sc["he"], sc["te"], sc["kernel_wt"], sc["param"],
                                                            sc["sza"], sc["raa"], sc["vza"], sc["flux"], M, 0)
    assert rc == 0
    nf, st = VP.last_diag()
    got = dict(radiance=I, reflectance=R, Q=Q, U=U, nfourier=nf.reshape(nobs, 2))
    rep = parity.parity_report("modis", 3, N, False, sc, got)
    print("parity modis_cloud", rep)
    assert rep["n"] + rep["n_special"] == 2 * nobs and rep["n"] >= 10 and rep["pass"], rep
    assert (I[3] == M).all()


def test_scalar_cloud_variants(mods):
    """VLIDORT_Scalar_GissCX_Cloud / VLIDORT_Scalar_LandMODIS_Cloud -> VLIDORT_Run_Scalar_Cloud (V283/VLIDORT_ScatMod.F90:2596): nstokes 1,
    P11 of aerosol and cloud matrices, the g / gI / gL arguments do not enter"""
    VP, vg, parity = mods
    nobs, N = 9, 8
    sc = _cloud_scene(vg, nobs, (550.0,), 64, 503)
    z = np.zeros_like(sc["tau"])
    cl = (sc["tauI"], sc["ssaI"], z, sc["pmomI"], sc["tauL"], sc["ssaL"], z + 0.3, sc["pmomL"])
    I, R, BR, rc = VP.scalar_gisscx_cloud(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["g"], sc["pmom"], *cl, sc["pe"],
                                          sc["he"], sc["te"], sc["u10m"], sc["v10m"], sc["mr"], sc["sza"], sc["raa"], sc["vza"], M, 0)
    assert rc == 0
    nf, st = VP.last_diag()
    assert (st == 0).all()
    flat = dict(sc); del flat["alpha"], flat["flux"]
    rep = parity.parity_report("gisscx", 1, N, False, flat, dict(radiance=I, reflectance=R, Q=None, U=None, nfourier=nf.reshape(nobs, 1)),
                               scalar_glue=True, tol_conv=5e-9)
    print("parity scalar gisscx_cloud", rep)
    assert rep["n"] == nobs and rep["pass"], str(rep)
    I, R, BR, rc = VP.scalar_brdf_modis_cloud(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["alpha"], sc["tau"], sc["ssa"], sc["g"], sc["pmom"],
                                              *cl, sc["pe"], sc["he"], sc["te"], sc["kernel_wt"], sc["param"], sc["sza"], sc["raa"], sc["vza"],
                                              sc["flux"], M, 0)
    assert rc == 0
    nf, st = VP.last_diag()
    rep = parity.parity_report("modis", 1, N, False, sc, dict(radiance=I, reflectance=R, Q=None, U=None, nfourier=nf.reshape(nobs, 1)),
                               scalar_glue=True, tol_conv=5e-9)
    print("parity scalar modis_cloud", rep)
    assert rep["n"] + rep["n_special"] == nobs and rep["n"] >= 5 and rep["pass"], str(rep)


def test_lambert_and_cx_cloud_wrappers_ignore_the_clouds_like_the_reference(mods):
    """VLIDORT_Vector_Lambert_Cloud / VLIDORT_Vector_CX_Cloud (f2py VECTOR_LAMBERT_CLOUD, VECTOR_CX_CLOUD) accept the cloud arrays and then
    call the clear-sky runner (GV/VLIDORT_LAMBERT.F90:843, GV/VLIDORT_BRDF_CX.F90:1118): the drop-in must equal the clear-sky entry bit for
    bit whatever the cloud arrays hold"""
    VP, vg, _ = mods
    nobs, N = 7, 6
    sc = _cloud_scene(vg, nobs, (550.0,), 40, 504)
    cl = [sc[k] for k in ("tauI", "ssaI", "pmomI", "tauL", "ssaL", "pmomL")]
    a = VP.vector_lambert_cloud(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["alpha"], sc["tau"], sc["ssa"], sc["pmom"], *cl, sc["pe"], sc["he"],
                                sc["te"], sc["albedo"], sc["sza"], sc["raa"], sc["vza"], sc["flux"], M, 0)
    b = VP.vector_lambert(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["alpha"], sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"], sc["te"],
                          sc["albedo"], sc["sza"], sc["raa"], sc["vza"], sc["flux"], M, 0)
    assert a[-1] == b[-1] == 0 and all(np.array_equal(x, y) for x, y in zip(a[:-1], b[:-1]))
    a = VP.vector_cx_cloud(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], *([None] * 6), sc["pe"], sc["he"],
                           sc["te"], sc["u10m"], sc["v10m"], sc["mr"], sc["sza"], sc["raa"], sc["vza"], M, 0)
    b = VP.vector_cx(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"], sc["te"], sc["u10m"],
                     sc["v10m"], sc["mr"], sc["sza"], sc["raa"], sc["vza"], M, 0)
    assert a[-1] == b[-1] == 0 and all(np.array_equal(x, y) for x, y in zip(a[:-1], b[:-1]))
