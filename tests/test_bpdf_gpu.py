"""BPDF surface: VLIDORT_Vector_BPDF (GV/VLIDORT_BRDF_MODIS_BPDF.F90:459; f2py VECTOR_BPDF) -- one BPDF-NDVI kernel
(BPDFNDVI_VFUNCTION + FRESNEL_VECTOR, VB/vbrdf_sup_kernels.f90:3017-3198,3560-3702; set-up VLIDORT_BPDF, V283/VLIDORT_SurfaceMod.F90:2279) --
through the C ABI against the oracle's restatement of the same Fortran (PARITY UNPINNED: the reference's saved VBRDF results do not include
this kernel).  Bar and protocol as everywhere: 1e-8 on I, reflectance, Q, U (relative to I), identical Fourier counts; the exact direct-bounce
elements BR, BR_Q, BR_U to 1e-12."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
M = 1e15


def test_vector_bpdf(built):
    from aeroapps_b200 import VLIDORT_POLAR_ as VP
    from oracle import vlidort_glue as vg, parity
    nobs, N = 16, 8
    sc = vg.synthetic_scene(nobs, nch_nm=(470.0, 865.0), nmom=80, seed=701)
    rng = np.random.default_rng(702)
    bp = np.zeros((3, 2, nobs))
    bp[0] = 1.5; bp[1] = rng.uniform(-0.1, 0.9, (2, nobs)); bp[2] = rng.uniform(2.0, 7.0, (2, nobs))
    bp[1, 0, 3] = 1.7                                             # NDVI outside [-1, 1] is replaced by 0 inside the kernel
    sc["vza"][5] = 0.0; sc["sza"][6] = M
    sc["param"] = np.asfortranarray(bp); sc.pop("kernel_wt", None)
    I, R, BR, Q, U, BQ, BU, rc = VP.vector_bpdf(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["pe"],
                                                sc["he"], sc["te"], sc["param"], sc["sza"], sc["raa"], sc["vza"], M, 0)
    assert rc == 0
    nf, st = VP.last_diag()
    got = dict(radiance=I, reflectance=R, Q=Q, U=U, nfourier=nf.reshape(nobs, 2))
    rep = parity.parity_report("bpdf", 3, N, False, sc, got)
    print("parity bpdf", rep)
    assert rep["n"] + rep["n_special"] == 2 * nobs and rep["n"] >= 2 * nobs - 2 and rep["pass"], str(rep)
    ref = parity.oracle_parallel("bpdf", 3, N, False, sc)
    ok = ref["status"] == 0
    for a, b in ((BR, ref["BR"]), (BQ, ref["BR_Q"]), (BU, ref["BR_U"])):
        assert np.all(np.abs(a[ok] - b[ok]) <= 1e-12 * np.abs(b[ok]) + 1e-15)
    assert np.abs(Q[ok]).max() > 0 and np.abs(BQ[ok]).max() > 0     # the surface polarises
    assert (I[6] == M).all()


def _bpdf_scene(nobs, seed):
    from oracle import vlidort_glue as vg
    sc = vg.synthetic_scene(nobs, nch_nm=(470.0, 865.0), nmom=80, seed=seed)
    rng = np.random.default_rng(seed + 1)
    bp = np.zeros((3, 2, nobs))
    bp[0] = 1.5; bp[1] = rng.uniform(-0.1, 0.9, (2, nobs)); bp[2] = rng.uniform(2.0, 7.0, (2, nobs))
    sc["bpdf"] = np.asfortranarray(bp)
    return sc


def _report(parity, kind, sc, N, outs, nf, nobs):
    I, R, BR, Q, U, BQ, BU, rc = outs
    assert rc == 0
    rep = parity.parity_report(kind, 3, N, False, sc, dict(radiance=I, reflectance=R, Q=Q, U=U, nfourier=nf.reshape(nobs, 2)))
    print("parity", kind, rep)
    ref = parity.oracle_parallel(kind, 3, N, False, sc)
    ok = ref["status"] == 0
    for a, b in ((BR, ref["BR"]), (BQ, ref["BR_Q"]), (BU, ref["BR_U"])):
        assert np.all(np.abs(a[ok] - b[ok]) <= 1e-12 * np.abs(b[ok]) + 1e-15)
    return rep, ok


def test_vector_brdf_modis_bpdf(built):
    """VLIDORT_Vector_LandMODIS_BPDF (GV/VLIDORT_BRDF_MODIS_BPDF.F90:23; f2py VECTOR_BRDF_MODIS_BPDF): Lambertian + Ross-thick + Li-sparse
    (element (1,1) only) mixed with the BPDF-NDVI Mueller block"""
    from aeroapps_b200 import VLIDORT_POLAR_ as VP
    from oracle import parity
    nobs, N = 14, 8
    sc = _bpdf_scene(nobs, 711)
    outs = VP.vector_brdf_modis_bpdf(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"], sc["te"],
                                     sc["kernel_wt"], sc["param"], sc["bpdf"], sc["sza"], sc["raa"], sc["vza"], M, 0)
    nf, st = VP.last_diag()
    rep, ok = _report(parity, "modis_bpdf", sc, N, outs, nf, nobs)
    assert rep["n"] + rep["n_special"] == 2 * nobs and rep["n"] >= nobs and rep["pass"], str(rep)
    assert np.abs(outs[5][ok]).max() > 0                             # BR_Q comes from the BPDF kernel


def test_vector_lambert_bpdf(built):
    """VLIDORT_Vector_Lambert_BPDF (GV/VLIDORT_LAMBERT_BPDF.F90:23; f2py VECTOR_LAMBERT_BPDF); a MISSING albedo makes the pixel MISSING"""
    from aeroapps_b200 import VLIDORT_POLAR_ as VP
    from oracle import parity
    nobs, N = 12, 6
    sc = _bpdf_scene(nobs, 721)
    sc["albedo"][4, 1] = M
    sc.pop("kernel_wt", None); sc.pop("param", None)
    outs = VP.vector_lambert_bpdf(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"], sc["te"],
                                  sc["albedo"], sc["bpdf"], sc["sza"], sc["raa"], sc["vza"], M, 0)
    nf, st = VP.last_diag()
    rep, ok = _report(parity, "lambert_bpdf", sc, N, outs, nf, nobs)
    assert rep["n"] == 2 * nobs - 1 and rep["n_special"] == 1 and rep["pass"], str(rep)
    assert outs[0][4, 1] == M and outs[0][4, 0] != M
