"""The remaining cloud forms of the reference's Fortran modules (none has an f2py wrapper) through the C ABI:

* cloud runner + new surface / geometry combinations, against the oracle (bar and protocol of oracle/parity.py: 1e-8, identical Fourier
  counts): VLIDORT_Vector_OCIGissCX_Cloud (GV/VLIDORT_BRDF_CX.F90:2468), VLIDORT_Scalar_OCIGissCX_Cloud (:2084),
  VLIDORT_Vector_LandMODIS_Cloud_MultiGeom (GV/VLIDORT_BRDF_MODIS.F90:1055);
* "place holder" wrappers that take the cloud arrays and call the CLEAR-SKY runner: VLIDORT_Scalar_Lambert_Cloud (GV/VLIDORT_LAMBERT.F90:524,
  adds alpha and the solar flux to the scalar Lambertian form), VLIDORT_Vector_Lambert_Cloud_MultiGeom (:881), VLIDORT_Vector_CX_Cloud_MultiGeom
  (GV/VLIDORT_BRDF_CX.F90:1147), VLIDORT_Vector_LandMODIS_BPDF_Cloud (GV/VLIDORT_BRDF_MODIS_BPDF.F90:232, adds alpha and the solar flux),
  VLIDORT_Vector_Lambert_BPDF_Cloud (GV/VLIDORT_LAMBERT_BPDF.F90:228): the cloud arrays must not matter (None is accepted) and the result
  must equal the oracle / the clear-sky entry."""
import numpy as np
import pytest

from .test_cloud_gpu import _cloud_scene

pytestmark = pytest.mark.gpu
M = 1e15
CL = ("tauI", "ssaI", "pmomI", "tauL", "ssaL", "pmomL")


@pytest.fixture(scope="module")
def mods(built):
    from aeroapps_b200 import VLIDORT_POLAR_ as VP
    from oracle import vlidort_glue as vg, parity
    return VP, vg, parity


def test_vector_ocigisscx_cloud(mods):
    VP, vg, parity = mods
    nobs, N = 10, 8
    sc = _cloud_scene(vg, nobs, (440.0, 865.0), 80, 601)
    sc["u10m"][2] = M                                                          # a MISSING wind
    cl = [sc[k] for k in CL]
    I, R, BR, Q, U, BQ, BU, rc = VP.vector_ocigisscx_cloud(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["alpha"], sc["tau"], sc["ssa"],
                                                           sc["pmom"], *cl, sc["pe"], sc["he"], sc["te"], sc["u10m"], sc["v10m"], sc["mr"],
                                                           sc["sza"], sc["raa"], sc["vza"], sc["flux"], M, 0)
    assert rc == 0
    nf, st = VP.last_diag()
    got = dict(radiance=I, reflectance=R, Q=Q, U=U, nfourier=nf.reshape(nobs, 2))
    rep = parity.parity_report("ocigisscx", 3, N, False, sc, got, tol_conv=5e-9)       # bar of the cloudy ocean scenes: see test_cloud_gpu.py
    print("parity ocigisscx_cloud", rep)
    assert rep["n"] + rep["n_special"] == 2 * nobs and rep["n"] == 2 * (nobs - 1) and rep["pass"], str(rep)
    assert (I[2] == M).all()
    ref = parity.oracle_parallel("ocigisscx", 3, N, False, sc)
    ok = ref["status"] == 0
    for a, b in ((BR, ref["BR"]), (BQ, ref["BR_Q"]), (BU, ref["BR_U"])):
        assert np.all(np.abs(a[ok] - b[ok]) <= 1e-12 * np.abs(b[ok]) + 1e-15)
    # the clouds, the trace gases and the flux all matter
    one = np.ones_like(sc["flux"]); z = np.zeros_like(sc["alpha"])
    I0 = VP.vector_ocigisscx(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["alpha"], sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"],
                             sc["te"], sc["u10m"], sc["v10m"], sc["mr"], sc["sza"], sc["raa"], sc["vza"], sc["flux"], M, 0)[0]
    assert np.max(np.abs(I[ok] - I0[ok]) / I0[ok]) > 0.05
    I1 = VP.vector_ocigisscx_cloud(sc["channels"], N, False, sc["ROT"], sc["depol"], z, sc["tau"], sc["ssa"], sc["pmom"], *cl, sc["pe"], sc["he"],
                                   sc["te"], sc["u10m"], sc["v10m"], sc["mr"], sc["sza"], sc["raa"], sc["vza"], one, M, 0)[0]
    assert np.max(np.abs(I[ok] - I1[ok]) / I1[ok]) > 1e-3


def test_scalar_ocigisscx_cloud(mods):
    VP, vg, parity = mods
    nobs, N = 9, 8
    sc = _cloud_scene(vg, nobs, (550.0,), 64, 602)
    z = np.zeros_like(sc["tau"])
    cl = (sc["tauI"], sc["ssaI"], z, sc["pmomI"], sc["tauL"], sc["ssaL"], z + 0.3, sc["pmomL"])
    I, R, BR, rc = VP.scalar_ocigisscx_cloud(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["alpha"], sc["tau"], sc["ssa"], sc["g"], sc["pmom"],
                                             *cl, sc["pe"], sc["he"], sc["te"], sc["u10m"], sc["v10m"], sc["mr"], sc["sza"], sc["raa"], sc["vza"],
                                             sc["flux"], M, 0)
    assert rc == 0
    nf, st = VP.last_diag()
    assert (st == 0).all()
    rep = parity.parity_report("ocigisscx", 1, N, False, sc, dict(radiance=I, reflectance=R, Q=None, U=None, nfourier=nf.reshape(nobs, 1)),
                               scalar_glue=True, tol_conv=5e-9)
    print("parity scalar ocigisscx_cloud", rep)
    assert rep["n"] == nobs and rep["pass"], str(rep)


def test_vector_brdf_modis_cloud_multigeom(mods):
    """every view of the ONE multi-geometry call equals the single-geometry oracle run of that view with the same clouds"""
    VP, vg, parity = mods
    nobs, N, G = 6, 6, 5
    sc = _cloud_scene(vg, nobs, (470.0, 660.0), 60, 603)
    nch = 2
    rng = np.random.default_rng(604)
    vza = np.asfortranarray(rng.uniform(0.0, 60.0, (nobs, G))); raa = np.asfortranarray(rng.uniform(0.0, 180.0, (nobs, G)))
    sza = np.asfortranarray(np.repeat(sc["sza"][:, None], G, axis=1))
    vza[3, 0] = M                                                              # first geometry MISSING: the whole observation is
    cl = [sc[k] for k in CL]
    I, R, BR, Q, U, BQ, BU, rc = VP.vector_brdf_modis_cloud_multigeom(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["alpha"], sc["tau"],
                                                                      sc["ssa"], sc["pmom"], *cl, sc["pe"], sc["he"], sc["te"], sc["kernel_wt"],
                                                                      sc["param"], sza, raa, vza, sc["flux"], M, 0)
    assert rc == 0 and I.shape == (nobs, nch, G)
    nf, st = VP.last_diag()
    assert (I[3] == M).all() and (BQ == 0).all()
    keep = [j for j in range(nobs) if j != 3]
    rep = lambda a, ax: np.asfortranarray(np.concatenate([np.take(a, keep, axis=ax)] * G, axis=ax))
    flat = dict(ROT=rep(sc["ROT"], 1), depol=sc["depol"], channels=sc["channels"], tau=rep(sc["tau"], 2), ssa=rep(sc["ssa"], 2),
                pmom=rep(sc["pmom"], 2), he=rep(sc["he"], 1), alpha=rep(sc["alpha"], 1), flux=rep(sc["flux"], 1),
                kernel_wt=rep(sc["kernel_wt"], 2), param=rep(sc["param"], 2),
                sza=sza[keep].ravel(order="F"), raa=raa[keep].ravel(order="F"), vza=vza[keep].ravel(order="F"))
    for k in CL:
        flat[k] = rep(sc[k], 2)
    nk = len(keep)
    fl = lambda a: np.asfortranarray(np.concatenate([a[keep, :, g] for g in range(G)], axis=0))      # (nobs,nch,G) -> (G*nk, nch), index g*nk + j
    got = dict(radiance=fl(I), reflectance=fl(R), Q=fl(Q), U=fl(U), nfourier=fl(nf.reshape(nobs, nch, G)))
    rp = parity.parity_report("modis", 3, N, False, flat, got, tol_conv=5e-9)
    print("parity modis_cloud_multigeom", rp)
    assert rp["n"] + rp["n_special"] == nk * G * nch and rp["n"] >= nk * G and rp["pass"], str(rp)
    ref = parity.oracle_parallel("modis", 3, N, False, flat)
    ok = ref["status"] == 0
    assert np.all(np.abs(fl(BR)[ok] - ref["BR"][ok]) <= 1e-12 * np.abs(ref["BR"][ok]) + 1e-15)


def test_scalar_lambert_cloud(mods):
    VP, vg, parity = mods
    nobs, N = 12, 8
    sc = _cloud_scene(vg, nobs, (354.0, 550.0), 48, 605)
    sc["albedo"][4, 1] = M
    cl = lambda: (sc["tauI"], sc["ssaI"], sc["g"], sc["pmomI"], sc["tauL"], sc["ssaL"], sc["g"], sc["pmomL"])
    call = lambda c: VP.scalar_lambert_cloud(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["alpha"], sc["tau"], sc["ssa"], sc["g"], sc["pmom"],
                                             *c, sc["pe"], sc["he"], sc["te"], sc["albedo"], sc["sza"], sc["raa"], sc["vza"], sc["flux"], M, 0)
    I, R, rc = call(cl())
    assert rc == 0
    nf, st = VP.last_diag()
    clear = {k: v for k, v in sc.items() if k not in CL}                       # the reference's runner is the clear-sky one
    rep = parity.parity_report("lambert", 1, N, False, clear, dict(radiance=I, reflectance=R, Q=None, U=None, nfourier=nf.reshape(nobs, 2)),
                               scalar_glue=True)
    print("parity scalar lambert_cloud", rep)
    assert rep["n"] + rep["n_special"] == 2 * nobs and rep["n"] == 2 * nobs - 1 and rep["pass"], str(rep)
    assert I[4, 1] == M
    I2, R2, rc = call([None] * 8)
    assert rc == 0 and np.array_equal(I, I2) and np.array_equal(R, R2)
    # alpha = 0 and unit flux: the plain scalar Lambertian entry
    I3 = VP.scalar_lambert_cloud(sc["channels"], N, False, sc["ROT"], sc["depol"], np.zeros_like(sc["alpha"]), sc["tau"], sc["ssa"], sc["g"],
                                 sc["pmom"], *([None] * 8), sc["pe"], sc["he"], sc["te"], sc["albedo"], sc["sza"], sc["raa"], sc["vza"],
                                 np.ones_like(sc["flux"]), M, 0)[0]
    I4 = VP.scalar_lambert(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["g"], sc["pmom"], sc["pe"], sc["he"], sc["te"],
                           sc["albedo"], sc["sza"], sc["raa"], sc["vza"], M, 0)[0]
    assert np.array_equal(I3, I4)


def test_clear_sky_placeholders_equal_their_clear_sky_entries(mods):
    VP, vg, parity = mods
    nobs, N, G = 5, 6, 3
    sc = _cloud_scene(vg, nobs, (470.0, 865.0), 40, 606)
    nch = 2
    cl = [sc[k] for k in CL]
    rng = np.random.default_rng(607)
    vza = np.asfortranarray(rng.uniform(0.0, 60.0, (nobs, G))); raa = np.asfortranarray(rng.uniform(0.0, 180.0, (nobs, G)))
    sza = np.asfortranarray(np.repeat(sc["sza"][:, None], G, axis=1))
    same = lambda a, b: a[-1] == b[-1] == 0 and all(np.array_equal(x, y) for x, y in zip(a[:-1], b[:-1]))
    # Lambert, multi-geometry: alpha enters, the flux stays 1
    a = VP.vector_lambert_cloud_multigeom(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["alpha"], sc["tau"], sc["ssa"], sc["pmom"], *cl,
                                          sc["pe"], sc["he"], sc["te"], sc["albedo"], sza, raa, vza, M, 0)
    b = VP.vector_lambert_multigeom(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["alpha"], sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"],
                                    sc["te"], sc["albedo"], sza, raa, vza, np.ones((nch, nobs), order="F"), M, 0)
    assert a[0].shape == (nobs, nch, G) and same(a, b)
    # Cox-Munk, multi-geometry
    a = VP.vector_cx_cloud_multigeom(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], *([None] * 6), sc["pe"],
                                     sc["he"], sc["te"], sc["u10m"], sc["v10m"], sc["mr"], sza, raa, vza, M, 0)
    b = VP.vector_cx_multigeom(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"], sc["te"],
                               sc["u10m"], sc["v10m"], sc["mr"], sza, raa, vza, M, 0)
    assert same(a, b)
    # Lambert + BPDF
    bp = np.zeros((3, nch, nobs)); bp[0] = 1.5; bp[1] = rng.uniform(-0.1, 0.9, (nch, nobs)); bp[2] = rng.uniform(2.0, 7.0, (nch, nobs))
    bp = np.asfortranarray(bp)
    a = VP.vector_lambert_bpdf_cloud(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], *cl, sc["pe"], sc["he"],
                                     sc["te"], sc["albedo"], bp, sc["sza"], sc["raa"], sc["vza"], M, 0)
    b = VP.vector_lambert_bpdf(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"], sc["te"],
                               sc["albedo"], bp, sc["sza"], sc["raa"], sc["vza"], M, 0)
    assert same(a, b)
    # MODIS RTLS + BPDF: alpha and the flux enter -> the oracle with the same glue; alpha = 0, flux = 1 -> the clear-sky entry
    sc["bpdf"] = bp
    call = lambda al, fx, c: VP.vector_brdf_modis_bpdf_cloud(sc["channels"], N, False, sc["ROT"], sc["depol"], al, sc["tau"], sc["ssa"], sc["pmom"], *c,
                                                             sc["pe"], sc["he"], sc["te"], sc["kernel_wt"], sc["param"], bp, sc["sza"], sc["raa"],
                                                             sc["vza"], fx, M, 0)
    I, R, BR, Q, U, BQ, BU, rc = call(sc["alpha"], sc["flux"], cl)
    assert rc == 0
    nf, st = VP.last_diag()
    clear = {k: v for k, v in sc.items() if k not in CL}
    rep = parity.parity_report("modis_bpdf", 3, N, False, clear, dict(radiance=I, reflectance=R, Q=Q, U=U, nfourier=nf.reshape(nobs, nch)))
    print("parity modis_bpdf_cloud", rep)
    assert rep["n"] + rep["n_special"] == nch * nobs and rep["n"] >= nobs and rep["pass"], str(rep)
    a = call(np.zeros_like(sc["alpha"]), np.ones_like(sc["flux"]), [None] * 6)
    b = VP.vector_brdf_modis_bpdf(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"], sc["te"],
                                  sc["kernel_wt"], sc["param"], bp, sc["sza"], sc["raa"], sc["vza"], M, 0)
    assert same(a, b)


def test_nobm_forms_without_adjustment_equal_their_siblings(mods):
    """NOBM water-leaving forms (GV/VLIDORT_BRDF_CX.F90:1351,2671,2883,3102,3320; f2py VECTOR_OCIGissCX_NOBM_Cloud).  With sleave_adjust False
    VLIDORT 2.8.3 multiplies every water-leaving term by TRANS_ATMOS_FINAL, which stays at the zero it is reset to unless the transmittance
    iteration or DO_EXTERNAL_WLEAVE is on (VM/vlidort_bvproblem.f90:1500-1513,1563-1571; VM/vlidort_masters_V2p8p3.f90:2493-2499; AeroApps sets
    neither, V283/VLIDORT_ScatMod.F90:388-405): the reference's answer is the one without water-leaving, ADJUSTED_SLEAVE = 0.  The entries must
    therefore equal their siblings (which the tests above pin to the oracle) bit for bit, whatever sleave / sleave_iso are, and must refuse
    sleave_adjust True instead of returning an unadjusted answer"""
    VP, vg, parity = mods
    nobs, N = 6, 6
    sc = _cloud_scene(vg, nobs, (440.0, 865.0), 40, 608)
    sc["v10m"][1] = M
    cl = [sc[k] for k in CL]
    scl = (sc["tauI"], sc["ssaI"], sc["g"], sc["pmomI"], sc["tauL"], sc["ssaL"], sc["g"], sc["pmomL"])
    sleave = np.asfortranarray(np.random.default_rng(609).uniform(0.0, 0.02, (2, nobs)))
    geo = (sc["sza"], sc["raa"], sc["vza"])
    atm = (sc["channels"], N, False, sc["ROT"], sc["depol"])
    lev = (sc["pe"], sc["he"], sc["te"], sc["u10m"], sc["v10m"], sc["mr"])
    same = lambda a, b: a[-1] == b[-1] == 0 and all(np.array_equal(x, y) for x, y in zip(a[:-1], b[:-1]))
    for iso in (True, False):
        a = VP.vector_ocigisscx_nobm_cloud(*atm, sc["alpha"], sc["tau"], sc["ssa"], sc["pmom"], *cl, *lev, sleave, iso, False, *geo, sc["flux"], M, 0)
        b = VP.vector_ocigisscx_cloud(*atm, sc["alpha"], sc["tau"], sc["ssa"], sc["pmom"], *cl, *lev, *geo, sc["flux"], M, 0)
        assert len(a) == 9 and same(a[:8], b) and a[8].shape == (nobs, 2) and (a[8] == 0).all()
        assert (a[0][1] == M).all() and (a[0][0] != M).all() and (a[0][0] > 0).all()
        a = VP.vector_gisscx_nobm_cloud(*atm, sc["tau"], sc["ssa"], sc["pmom"], *cl, *lev, sleave, iso, False, *geo, M, 0)
        b = VP.vector_gisscx_cloud(*atm, sc["tau"], sc["ssa"], sc["pmom"], *cl, *lev, *geo, M, 0)
        assert same(a[:8], b) and (a[8] == 0).all()
        a = VP.vector_cx_nobm(*atm, sc["tau"], sc["ssa"], sc["pmom"], *lev, sleave, iso, False, *geo, M, 0)
        b = VP.vector_cx(*atm, sc["tau"], sc["ssa"], sc["pmom"], *lev, *geo, M, 0)
        assert same(a[:8], b) and (a[8] == 0).all()
        a = VP.scalar_gisscx_nobm_cloud(*atm, sc["tau"], sc["ssa"], sc["g"], sc["pmom"], *scl, *lev, sleave, iso, False, *geo, M, 0)
        b = VP.scalar_gisscx_cloud(*atm, sc["tau"], sc["ssa"], sc["g"], sc["pmom"], *scl, *lev, *geo, M, 0)
        assert same(a[:4], b) and (a[4] == 0).all()
        a = VP.scalar_ocigisscx_nobm_cloud(*atm, sc["alpha"], sc["tau"], sc["ssa"], sc["g"], sc["pmom"], *scl, *lev, sleave, iso, False, *geo, sc["flux"],
                                           M, 0)
        b = VP.scalar_ocigisscx_cloud(*atm, sc["alpha"], sc["tau"], sc["ssa"], sc["g"], sc["pmom"], *scl, *lev, *geo, sc["flux"], M, 0)
        assert same(a[:4], b) and (a[4] == 0).all()
    with pytest.raises(RuntimeError, match="sleave_adjust"):
        VP.vector_ocigisscx_nobm_cloud(*atm, sc["alpha"], sc["tau"], sc["ssa"], sc["pmom"], *cl, *lev, sleave, True, True, *geo, sc["flux"], M, 0)
    with pytest.raises(RuntimeError, match="sleave_adjust"):
        VP.vector_cx_nobm(*atm, sc["tau"], sc["ssa"], sc["pmom"], *lev, sleave, True, True, *geo, M, 0)
