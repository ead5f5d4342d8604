"""Pins oracle/vbrdf_oracle.cpp to the reference's own saved BRDF self-test (GISS Cox-Munk block): exact
direct-bounce BRDF at three azimuths and the Fourier components m=0,1 of all four BRDF arrays, to the 8
significant digits the reference prints."""
import ctypes as C
import os
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def orc(built):
    from oracle import vlidort_glue as vg
    return vg


def test_giss_coxmunk_self_tester(orc):
    g = np.load(os.path.join(HERE, "golden", "vbrdf_self_tester_giss.npz"))
    sigma2 = 0.003 + 0.00512 * 15.0
    refrac = float(np.float32(1.335) + np.float32(0.006) * (np.float32(34.3) / np.float32(34.3)))
    pars = [0.5 * sigma2, refrac, 0.0, 0.0]
    qidx = [[4 * a + b for b in range(3)] for a in range(3)]
    for ia, azm in enumerate((0.0, 79.0, 180.0)):
        out = orc.vbrdf(3, 1, [10], [1.0], [pars], [0], 100, True, 35.0, 10.0, azm)
        assert out["status"] == 0
        ex = out["exactdb"][qidx]
        ref = g["exact"][ia]
        assert np.allclose(ex, ref, rtol=2e-7, atol=1e-14), (azm, ex, ref)
    for m in range(2):
        got = [out["brdf_f"][m][:, 0, 0][qidx], out["brdf_f_0"][m][:, 0][qidx], out["user_brdf_f"][m][:, 0][qidx],
               out["user_brdf_f_0"][m][qidx]]
        for k in range(4):
            assert np.allclose(got[k], g["fourier"][m, k], rtol=2e-7, atol=1e-14), (m, k, got[k], g["fourier"][m, k])


def test_three_kernel_supcheck_obsgeo(orc):
    """scalar Cox-Munk (the kernel AeroApps' CX entry points use) + two MODIS-type kernels, half-range azimuth
    quadrature, multi-kernel accumulation: VLIDORT's saved results_brdf_supcheck_obsgeo.res, all 16 Fourier terms."""
    g = np.load(os.path.join(HERE, "golden", "vbrdf_supcheck_obsgeo.npz"))
    geoms = [(35.0, 10.0, 10.0), (67.0, 20.0, 90.0), (75.0, 40.0, 170.0)]
    pars = [[0, 0, 0, 0], [2.0, 1.0, 0, 0], [0.0798, 1.779556, 0, 0]]
    for ib, (sza, vza, raa) in enumerate(geoms):
        out = orc.vbrdf(1, 8, [2, 5, 9], [0.3, 0.1, 0.1], pars, [0, 0, 0], 100, True, sza, vza, raa)
        assert out["status"] == 0
        assert abs(out["exactdb"][0] / g["exact"][ib] - 1) < 1e-10
        kw = dict(rtol=3e-7, atol=2e-7 * 1e-2)
        assert np.allclose(out["brdf_f"][:, 0], g["brdf_f"], **kw)
        assert np.allclose(out["brdf_f_0"][:, 0], g["brdf_f_0"][:, :, ib], **kw)
        assert np.allclose(out["user_brdf_f"][:, 0], g["user_brdf_f"][:, ib], **kw)
        assert np.allclose(out["user_brdf_f_0"][:, 0], g["user_brdf_f_0"][:, ib], **kw)


KPARS = [[0, 0, 0, 0], [2.0, 1.0, 0, 0], [0.0798, 1.779556, 0, 0]]


def _rt(orc, nstokes, sza, vza, raa, atm, nf=15):
    from oracle import vlidort_oracle as vo
    height, deltau, omega, greek = atm
    b = orc.vbrdf(1, 8, [2, 5, 9], [0.3, 0.1, 0.1], KPARS, [0, 0, 0], 100, True, sza, vza, raa)
    st, _, status, _, _ = vo.vlidort_master(nstokes, 8, height, deltau, omega, greek, sza, vza, raa, lambertian=False, brdf=b,
                                           focorr=True, focorr_nadir=True, deltam=True, accuracy=1e-4, max_fourier=nf,
                                           asymtx_tolerance=1e-20)
    assert status == 0
    return st


def test_rt_with_brdf_matches_brdfplus_tester(orc):
    """the BRDF surface inside the RT solve (BVP surface condition, post-processed reflection, exact direct bounce):
    VLIDORT's saved BRDF-plus tester results, 6 printed digits; scalar obs-geometry and vector RT + scalar BRDF (the
    combination AeroApps' Cox-Munk and MODIS entry points use).  In lattice mode the reference stops the Fourier
    series per beam once ALL its geometries have converged, so the count is searched per beam (as in
    test_vlidort_oracle_golden)."""
    import sys; sys.path.insert(0, HERE)
    from test_vlidort_oracle_golden import build_tester_atmosphere, GOLD
    z = np.load(GOLD); g = np.load(os.path.join(HERE, "golden", "vbrdfplus_toa.npz"))
    atm = build_tester_atmosphere(z, 1)
    for (sza, vza, raa), ref in zip([(35.0, 10.0, 10.0), (67.0, 20.0, 90.0), (75.0, 40.0, 170.0)], g["obsgeo_i"][0]):
        st = _rt(orc, 1, sza, vza, raa, atm)
        assert abs(st[0] - ref) < 0.6e-5 * abs(ref) + 1e-12
    atm = build_tester_atmosphere(z, 3)
    k = 0
    for sza in (35.0, 67.0, 75.0, 82.0):
        geoms = [(vza, raa) for vza in (10.0, 20.0, 40.0) for raa in (10.0, 90.0, 170.0)]
        found = None
        for nf in range(15, 2, -1):
            ok = True
            for j, (vza, raa) in enumerate(geoms):
                st = _rt(orc, 3, sza, vza, raa, atm, nf)
                for c in range(3):
                    ref = g["lattice_iqu"][c, k + j]
                    # half a unit of the 6th printed digit
                    if abs(st[c] - ref) > 0.5e-5 * 10.0 ** np.floor(np.log10(abs(ref))) * 1.02:
                        ok = False
            if ok:
                found = nf
                break
        assert found is not None, f"no Fourier count reproduces the saved BRDF-plus results for SZA={sza}"
        k += 9
