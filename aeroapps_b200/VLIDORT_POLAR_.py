"""Host-side mirror of the reference's f2py module ``VLIDORT_POLAR_`` (LV/VLIDORT_POLAR_py.F90, built by
LV/CMakeLists.txt:13-19): same lower-case function names, positional argument order (dimension integers are
inferred from the array shapes, as f2py does) and returned tuples.  Everything is computed by the CUDA library
through the C ABI of include/aeroapps_b200_rt.h; there is no CPU fallback.

    ROT, depol, rc                      = rot_calc(channels, pe, he, te, MISSING, verbose)
    I, reflectance, rc                  = scalar_lambert(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, g,
                                                         pmom, pe, he, te, albedo, sza, raa, vza, MISSING, verbose)
    I, reflectance, Q, U, rc            = vector_lambert(channels, nstreams, plane_parallel, ROT, depol, alpha, tau,
                                                         ssa, pmom, pe, he, te, albedo, sza, raa, vza, flux_factor,
                                                         MISSING, verbose)
"""
import ctypes as C
import numpy as np

from ._lib import lib, check

_dp = C.POINTER(C.c_double)


def _F(a):
    return np.asfortranarray(a, dtype=np.float64)


def _v(a):
    return np.ascontiguousarray(a, dtype=np.float64).reshape(-1)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _dims(tau, pmom):
    km, nch, nobs = tau.shape
    assert pmom.shape[:3] == (km, nch, nobs), "pmom must be (km,nch,nobs,nMom,nPol)"
    return km, nch, nobs, pmom.shape[3], pmom.shape[4]


def rot_calc(channels, pe, he, te, MISSING, verbose):
    pe, he, te = _F(pe), _F(he), _F(te)
    ch = _v(channels)
    km, nobs, nch = pe.shape[0] - 1, pe.shape[1], len(ch)
    ROT = np.zeros((km, nobs, nch), order="F"); depol = np.zeros(nch)
    rc = C.c_int(0)
    check(lib().ab200_rot_calc(km, nch, nobs, _p(ch), _p(pe), _p(he), _p(te), C.c_double(MISSING), int(verbose), _p(ROT),
                               _p(depol), C.byref(rc)), "ab200_rot_calc")
    return ROT, depol, rc.value


def scalar_lambert(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, g, pmom, pe, he, te, albedo,
                   solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose):
    ROT, tau, ssa, g, pmom, pe, he, te, albedo = (_F(x) for x in (ROT, tau, ssa, g, pmom, pe, he, te, albedo))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    rad = np.zeros((nobs, nch), order="F"); refl = np.zeros((nobs, nch), order="F")
    rc = C.c_int(0)
    check(lib().ab200_vlidort_scalar_lambert(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol,
                                             _p(ROT), _p(dp), _p(tau), _p(ssa), _p(g), _p(pmom), _p(pe), _p(he), _p(te),
                                             _p(albedo), _p(sza), _p(raa), _p(vza), C.c_double(MISSING), int(verbose),
                                             _p(rad), _p(refl), C.byref(rc)), "ab200_vlidort_scalar_lambert")
    return rad, refl, rc.value


def vector_lambert(channels, nstreams, plane_parallel, ROT, depol, alpha, tau, ssa, pmom, pe, he, te, albedo,
                   solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose):
    ROT, alpha, tau, ssa, pmom, pe, he, te, albedo, flux_factor = (_F(x) for x in (ROT, alpha, tau, ssa, pmom, pe, he, te,
                                                                                  albedo, flux_factor))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(4)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_lambert(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol,
                                             _p(ROT), _p(dp), _p(alpha), _p(tau), _p(ssa), _p(pmom), _p(pe), _p(he), _p(te),
                                             _p(albedo), _p(sza), _p(raa), _p(vza), _p(flux_factor), C.c_double(MISSING),
                                             int(verbose), *[_p(o) for o in outs], C.byref(rc)), "ab200_vlidort_vector_lambert")
    return outs[0], outs[1], outs[2], outs[3], rc.value


def _cx(entry, channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, pe, he, te, U10m, V10m, mr, solar_zenith,
        relat_azymuth, sensor_zenith, MISSING, verbose):
    ROT, tau, ssa, pmom, pe, he, te = (_F(x) for x in (ROT, tau, ssa, pmom, pe, he, te))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    u, v, m = _v(U10m), _v(V10m), _v(mr)
    assert len(u) == nobs and len(v) == nobs and len(m) == nch
    outs = [np.zeros((nobs, nch), order="F") for _ in range(7)]
    rc = C.c_int(0)
    check(getattr(lib(), entry)(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, _p(ROT), _p(dp),
                                _p(tau), _p(ssa), _p(pmom), _p(pe), _p(he), _p(te), _p(u), _p(v), _p(m), _p(sza), _p(raa),
                                _p(vza), C.c_double(MISSING), int(verbose), *[_p(o) for o in outs], C.byref(rc)), entry)
    return (*outs, rc.value)


def vector_cx(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, pe, he, te, U10m, V10m, mr, solar_zenith,
              relat_azymuth, sensor_zenith, MISSING, verbose):
    """f2py VECTOR_CX (VLIDORT_POLAR_py.F90:830): returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U, rc)."""
    return _cx("ab200_vlidort_vector_cx", channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, pe, he, te, U10m, V10m,
               mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose)


def vector_gisscx(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, pe, he, te, U10m, V10m, mr, solar_zenith,
                  relat_azymuth, sensor_zenith, MISSING, verbose):
    """f2py VECTOR_GissCX (VLIDORT_POLAR_py.F90:673): returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U, rc)."""
    return _cx("ab200_vlidort_vector_gisscx", channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, pe, he, te, U10m,
               V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose)


def vector_ocicx(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, pe, he, te, U10m, V10m, mr, solar_zenith,
                 relat_azymuth, sensor_zenith, MISSING, verbose):
    """f2py VECTOR_OCICX (VLIDORT_POLAR_py.F90:903): Cox-Munk with the OCI slope variance; returns (radiance, reflectance, BR,
    Q, U, BR_Q, BR_U, rc)."""
    return _cx("ab200_vlidort_vector_ocicx", channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, pe, he, te, U10m,
               V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose)


def vector_ocigisscx(channels, nstreams, plane_parallel, ROT, depol, alpha, tau, ssa, pmom, pe, he, te, U10m, V10m, mr,
                     solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose):
    """f2py VECTOR_OCIGissCX (VLIDORT_POLAR_py.F90:976): GISS Cox-Munk with the OCI slope variance, trace-gas absorption
    alpha (km,nobs,nch) and flux_factor (nch,nobs); returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U, rc)."""
    ROT, alpha, tau, ssa, pmom, pe, he, te, flux_factor = (_F(x) for x in (ROT, alpha, tau, ssa, pmom, pe, he, te, flux_factor))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    assert alpha.shape == ROT.shape and flux_factor.shape == (nch, nobs)
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    u, v, m = _v(U10m), _v(V10m), _v(mr)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(7)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_ocigisscx(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol,
                                               _p(ROT), _p(dp), _p(alpha), _p(tau), _p(ssa), _p(pmom), _p(pe), _p(he), _p(te),
                                               _p(u), _p(v), _p(m), _p(sza), _p(raa), _p(vza), _p(flux_factor),
                                               C.c_double(MISSING), int(verbose), *[_p(o) for o in outs], C.byref(rc)),
          "ab200_vlidort_vector_ocigisscx")
    return (*outs, rc.value)


def vector_brdf_modis(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, pe, he, te, kernel_wt, param,
                      solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose):
    """f2py VECTOR_BRDF_MODIS (VLIDORT_POLAR_py.F90:137): kernel_wt (3,nch,nobs) = fiso,fgeo,fvol; param (2,nch,nobs).
    Returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U, rc)."""
    ROT, tau, ssa, pmom, pe, he, te, kernel_wt, param = (_F(x) for x in (ROT, tau, ssa, pmom, pe, he, te, kernel_wt, param))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    assert kernel_wt.shape[1:] == (nch, nobs) and param.shape[1:] == (nch, nobs)
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(7)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_brdf_modis(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol,
                                                kernel_wt.shape[0], param.shape[0], _p(ROT), _p(dp), _p(tau), _p(ssa), _p(pmom),
                                                _p(pe), _p(he), _p(te), _p(kernel_wt), _p(param), _p(sza), _p(raa), _p(vza),
                                                C.c_double(MISSING), int(verbose), *[_p(o) for o in outs], C.byref(rc)),
          "ab200_vlidort_vector_brdf_modis")
    return (*outs, rc.value)


def vector_bpdf(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, pe, he, te, BPDFparam, solar_zenith, relat_azymuth,
                sensor_zenith, MISSING, verbose):
    """f2py VECTOR_BPDF (VLIDORT_POLAR_py.F90:295): BPDFparam (3,nch,nobs) = refractive index, NDVI, scaling factor.
    Returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U, rc)."""
    ROT, tau, ssa, pmom, pe, he, te, bp = (_F(x) for x in (ROT, tau, ssa, pmom, pe, he, te, BPDFparam))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    assert bp.shape == (3, nch, nobs), "BPDFparam must be (3,nch,nobs)"
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(7)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_bpdf(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, 3, _p(ROT), _p(dp), _p(tau),
                                          _p(ssa), _p(pmom), _p(pe), _p(he), _p(te), _p(bp), _p(sza), _p(raa), _p(vza), C.c_double(MISSING),
                                          int(verbose), *[_p(o) for o in outs], C.byref(rc)), "ab200_vlidort_vector_bpdf")
    return (*outs, rc.value)


def vector_brdf_modis_bpdf(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, pe, he, te, kernel_wt, RTLSparam, BPDFparam,
                           solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose):
    """f2py VECTOR_BRDF_MODIS_BPDF (VLIDORT_POLAR_py.F90:214): kernel_wt (3,nch,nobs), RTLSparam (2,nch,nobs), BPDFparam (3,nch,nobs).
    Returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U, rc)."""
    ROT, tau, ssa, pmom, pe, he, te, kw, rp, bp = (_F(x) for x in (ROT, tau, ssa, pmom, pe, he, te, kernel_wt, RTLSparam, BPDFparam))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    assert kw.shape == (3, nch, nobs) and rp.shape == (2, nch, nobs) and bp.shape == (3, nch, nobs)
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(7)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_brdf_modis_bpdf(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, 3, 2, 3, _p(ROT),
                                                     _p(dp), _p(tau), _p(ssa), _p(pmom), _p(pe), _p(he), _p(te), _p(kw), _p(rp), _p(bp), _p(sza),
                                                     _p(raa), _p(vza), C.c_double(MISSING), int(verbose), *[_p(o) for o in outs], C.byref(rc)),
          "ab200_vlidort_vector_brdf_modis_bpdf")
    return (*outs, rc.value)


def vector_lambert_bpdf(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, pe, he, te, albedo, BPDFparam, solar_zenith,
                        relat_azymuth, sensor_zenith, MISSING, verbose):
    """f2py VECTOR_LAMBERT_BPDF (VLIDORT_POLAR_py.F90:528): albedo (nobs,nch), BPDFparam (3,nch,nobs).
    Returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U, rc)."""
    ROT, tau, ssa, pmom, pe, he, te, alb, bp = (_F(x) for x in (ROT, tau, ssa, pmom, pe, he, te, albedo, BPDFparam))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    assert alb.shape == (nobs, nch) and bp.shape == (3, nch, nobs)
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(7)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_lambert_bpdf(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, 3, _p(ROT), _p(dp),
                                                  _p(tau), _p(ssa), _p(pmom), _p(pe), _p(he), _p(te), _p(alb), _p(bp), _p(sza), _p(raa), _p(vza),
                                                  C.c_double(MISSING), int(verbose), *[_p(o) for o in outs], C.byref(rc)),
          "ab200_vlidort_vector_lambert_bpdf")
    return (*outs, rc.value)


def scalar_cx(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, g, pmom, pe, he, te, U10m, V10m, mr, solar_zenith,
              relat_azymuth, sensor_zenith, MISSING, verbose):
    """VLIDORT_Scalar_CX (VLIDORT_BRDF_CX.F90:420): returns (radiance, reflectance, BRDF (1,nobs,nch), rc)."""
    ROT, tau, ssa, g, pmom, pe, he, te = (_F(x) for x in (ROT, tau, ssa, g, pmom, pe, he, te))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    u, v, m = _v(U10m), _v(V10m), _v(mr)
    rad = np.zeros((nobs, nch), order="F"); refl = np.zeros((nobs, nch), order="F"); brdf = np.zeros((1, nobs, nch), order="F")
    rc = C.c_int(0)
    check(lib().ab200_vlidort_scalar_cx(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, _p(ROT),
                                        _p(dp), _p(tau), _p(ssa), _p(g), _p(pmom), _p(pe), _p(he), _p(te), _p(u), _p(v), _p(m),
                                        _p(sza), _p(raa), _p(vza), C.c_double(MISSING), int(verbose), _p(rad), _p(refl),
                                        _p(brdf), C.byref(rc)), "ab200_vlidort_scalar_cx")
    return rad, refl, brdf, rc.value


def scalar_gisscx(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, g, pmom, pe, he, te, U10m, V10m, mr, solar_zenith,
                  relat_azymuth, sensor_zenith, MISSING, verbose):
    """VLIDORT_Scalar_GissCX (VLIDORT_BRDF_CX.F90:1553): returns (radiance, reflectance, BRDF (1,nobs,nch), rc)."""
    ROT, tau, ssa, g, pmom, pe, he, te = (_F(x) for x in (ROT, tau, ssa, g, pmom, pe, he, te))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    u, v, m = _v(U10m), _v(V10m), _v(mr)
    rad = np.zeros((nobs, nch), order="F"); refl = np.zeros((nobs, nch), order="F"); brdf = np.zeros((1, nobs, nch), order="F")
    rc = C.c_int(0)
    check(lib().ab200_vlidort_scalar_gisscx(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, _p(ROT),
                                            _p(dp), _p(tau), _p(ssa), _p(g), _p(pmom), _p(pe), _p(he), _p(te), _p(u), _p(v),
                                            _p(m), _p(sza), _p(raa), _p(vza), C.c_double(MISSING), int(verbose), _p(rad),
                                            _p(refl), _p(brdf), C.byref(rc)), "ab200_vlidort_scalar_gisscx")
    return rad, refl, brdf, rc.value


def scalar_brdf_modis(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, g, pmom, pe, he, te, kernel_wt, param,
                      solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose):
    """VLIDORT_Scalar_LandMODIS (VLIDORT_BRDF_MODIS.F90:35): returns (radiance, reflectance, BR, rc)."""
    ROT, tau, ssa, g, pmom, pe, he, te, kernel_wt, param = (_F(x) for x in (ROT, tau, ssa, g, pmom, pe, he, te, kernel_wt, param))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(3)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_scalar_brdf_modis(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol,
                                                _p(ROT), _p(dp), _p(tau), _p(ssa), _p(g), _p(pmom), _p(pe), _p(he), _p(te),
                                                _p(kernel_wt), _p(param), _p(sza), _p(raa), _p(vza), C.c_double(MISSING),
                                                int(verbose), *[_p(o) for o in outs], C.byref(rc)),
          "ab200_vlidort_scalar_brdf_modis")
    return (*outs, rc.value)


# ---- multi-geometry forms: the generic Fortran interfaces VLIDORT_Vector_Lambert / VLIDORT_Vector_CX / VLIDORT_Vector_LandMODIS over
#      SingleGeom / MultiGeom (GV/VLIDORT_LAMBERT.F90:14-17,346-520, GV/VLIDORT_BRDF_CX.F90:27-30,766-947, GV/VLIDORT_BRDF_MODIS.F90:413).
#      solar_zenith, relat_azymuth, sensor_zenith are (nobs,ngeom); outputs (nobs,nch,ngeom).  With one sun per observation (the 60 view
#      angles per column of LV/polar_vlidort.py:935-950) the eigenproblems and the boundary-value solve are shared by all views.
def _geom(solar_zenith, relat_azymuth, sensor_zenith, nobs):
    a = [np.asfortranarray(np.asarray(x, dtype=np.float64).reshape(nobs, -1)) for x in (solar_zenith, relat_azymuth, sensor_zenith)]
    ngeom = a[1].shape[1]
    if a[0].shape[1] == 1 and ngeom > 1:
        a[0] = np.asfortranarray(np.repeat(a[0], ngeom, axis=1))
    assert a[0].shape == a[1].shape == a[2].shape == (nobs, ngeom), "angles must be (nobs, ngeom)"
    return a[0], a[1], a[2], ngeom


def _opt(a):
    return None if a is None else _p(_F(a))


def vector_lambert_cloud(channels, nstreams, plane_parallel, ROT, depol, alpha, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te,
                         albedo, solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose):
    """f2py VECTOR_LAMBERT_CLOUD (VLIDORT_POLAR_py.F90:371).  The reference wrapper calls the clear-sky runner: the cloud arrays do not
    enter its result (GV/VLIDORT_LAMBERT.F90:843), nor this one (they may be None).  Returns (radiance, reflectance, Q, U, rc)."""
    ROT, alpha, tau, ssa, pmom, pe, he, te, albedo, flux_factor = (_F(x) for x in (ROT, alpha, tau, ssa, pmom, pe, he, te, albedo, flux_factor))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(4)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_lambert_cloud(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, _p(ROT), _p(dp),
                                                   _p(alpha), _p(tau), _p(ssa), _p(pmom), _opt(tauI), _opt(ssaI), _opt(pmomI), _opt(tauL),
                                                   _opt(ssaL), _opt(pmomL), _p(pe), _p(he), _p(te), _p(albedo), _p(sza), _p(raa), _p(vza),
                                                   _p(flux_factor), C.c_double(MISSING), int(verbose), *[_p(o) for o in outs], C.byref(rc)),
          "ab200_vlidort_vector_lambert_cloud")
    return outs[0], outs[1], outs[2], outs[3], rc.value


def vector_cx_cloud(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m, V10m,
                    mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose):
    """f2py VECTOR_CX_CLOUD (VLIDORT_POLAR_py.F90:746); cloud arrays ignored as in the reference (GV/VLIDORT_BRDF_CX.F90:1118).
    Returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U, rc)."""
    ROT, tau, ssa, pmom, pe, he, te = (_F(x) for x in (ROT, tau, ssa, pmom, pe, he, te))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    u, v, m = _v(U10m), _v(V10m), _v(mr)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(7)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_cx_cloud(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, _p(ROT), _p(dp), _p(tau),
                                              _p(ssa), _p(pmom), _opt(tauI), _opt(ssaI), _opt(pmomI), _opt(tauL), _opt(ssaL), _opt(pmomL), _p(pe),
                                              _p(he), _p(te), _p(u), _p(v), _p(m), _p(sza), _p(raa), _p(vza), C.c_double(MISSING), int(verbose),
                                              *[_p(o) for o in outs], C.byref(rc)), "ab200_vlidort_vector_cx_cloud")
    return (*outs, rc.value)


def vector_gisscx_cloud(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m,
                        V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose):
    """VLIDORT_Vector_GissCx_Cloud (GV/VLIDORT_BRDF_CX.F90:2276): aerosol + ice + liquid cloud layers over the GISS Cox-Munk ocean.
    Returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U, rc)."""
    ROT, tau, ssa, pmom, pe, he, te, tauI, ssaI, pmomI, tauL, ssaL, pmomL = (_F(x) for x in (ROT, tau, ssa, pmom, pe, he, te, tauI, ssaI, pmomI,
                                                                                              tauL, ssaL, pmomL))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    assert tauI.shape == tau.shape and tauL.shape == tau.shape and pmomI.shape == pmom.shape and pmomL.shape == pmom.shape
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    u, v, m = _v(U10m), _v(V10m), _v(mr)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(7)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_gisscx_cloud(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, _p(ROT), _p(dp),
                                                  _p(tau), _p(ssa), _p(pmom), _p(tauI), _p(ssaI), _p(pmomI), _p(tauL), _p(ssaL), _p(pmomL),
                                                  _p(pe), _p(he), _p(te), _p(u), _p(v), _p(m), _p(sza), _p(raa), _p(vza), C.c_double(MISSING),
                                                  int(verbose), *[_p(o) for o in outs], C.byref(rc)), "ab200_vlidort_vector_gisscx_cloud")
    return (*outs, rc.value)


def vector_brdf_modis_cloud(channels, nstreams, plane_parallel, ROT, depol, alpha, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he,
                            te, kernel_wt, param, solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose):
    """f2py VECTOR_BRDF_MODIS_CLOUD (VLIDORT_POLAR_py.F90:43): returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U, rc)."""
    ROT, alpha, tau, ssa, pmom, pe, he, te, tauI, ssaI, pmomI, tauL, ssaL, pmomL, kernel_wt, param, flux_factor = (
        _F(x) for x in (ROT, alpha, tau, ssa, pmom, pe, he, te, tauI, ssaI, pmomI, tauL, ssaL, pmomL, kernel_wt, param, flux_factor))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    assert tauI.shape == tau.shape and tauL.shape == tau.shape and pmomI.shape == pmom.shape and pmomL.shape == pmom.shape
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(7)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_brdf_modis_cloud(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol,
                                                      kernel_wt.shape[0], param.shape[0], _p(ROT), _p(dp), _p(alpha), _p(tau), _p(ssa), _p(pmom),
                                                      _p(tauI), _p(ssaI), _p(pmomI), _p(tauL), _p(ssaL), _p(pmomL), _p(pe), _p(he), _p(te),
                                                      _p(kernel_wt), _p(param), _p(sza), _p(raa), _p(vza), _p(flux_factor), C.c_double(MISSING),
                                                      int(verbose), *[_p(o) for o in outs], C.byref(rc)), "ab200_vlidort_vector_brdf_modis_cloud")
    return (*outs, rc.value)


def scalar_gisscx_cloud(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te,
                        U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose):
    """VLIDORT_Scalar_GissCX_Cloud (VLIDORT_BRDF_CX.F90:1898): returns (radiance, reflectance, BRDF (nobs,nch), rc)."""
    arrs = [_F(x) for x in (ROT, tau, ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te)]
    km, nch, nobs, nMom, nPol = _dims(arrs[1], arrs[4])
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    u, v, m = _v(U10m), _v(V10m), _v(mr)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(3)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_scalar_gisscx_cloud(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, _p(arrs[0]), _p(dp),
                                                  *[_p(a) for a in arrs[1:]], _p(u), _p(v), _p(m), _p(sza), _p(raa), _p(vza), C.c_double(MISSING),
                                                  int(verbose), *[_p(o) for o in outs], C.byref(rc)), "ab200_vlidort_scalar_gisscx_cloud")
    return (*outs, rc.value)


def scalar_brdf_modis_cloud(channels, nstreams, plane_parallel, ROT, depol, alpha, tau, ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL,
                            pe, he, te, kernel_wt, param, solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose):
    """VLIDORT_Scalar_LandMODIS_Cloud (VLIDORT_BRDF_MODIS.F90:625): returns (radiance, reflectance, BR (nobs,nch), rc)."""
    arrs = [_F(x) for x in (ROT, alpha, tau, ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te, kernel_wt, param)]
    km, nch, nobs, nMom, nPol = _dims(arrs[2], arrs[5])
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    ff = _F(flux_factor)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(3)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_scalar_brdf_modis_cloud(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, _p(arrs[0]),
                                                      _p(dp), *[_p(a) for a in arrs[1:]], _p(sza), _p(raa), _p(vza), _p(ff), C.c_double(MISSING),
                                                      int(verbose), *[_p(o) for o in outs], C.byref(rc)), "ab200_vlidort_scalar_brdf_modis_cloud")
    return (*outs, rc.value)


def vector_lambert_multigeom(channels, nstreams, plane_parallel, ROT, depol, alpha, tau, ssa, pmom, pe, he, te, albedo,
                             solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose):
    """returns (radiance, reflectance, Q, U (nobs,nch,ngeom), rc)"""
    ROT, alpha, tau, ssa, pmom, pe, he, te, albedo, flux_factor = (_F(x) for x in (ROT, alpha, tau, ssa, pmom, pe, he, te, albedo, flux_factor))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    sza, raa, vza, ngeom = _geom(solar_zenith, relat_azymuth, sensor_zenith, nobs)
    ch, dp = _v(channels), _v(depol)
    outs = [np.zeros((nobs, nch, ngeom), order="F") for _ in range(4)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_lambert_multigeom(km, nch, nobs, ngeom, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol,
                                                       _p(ROT), _p(dp), _p(alpha), _p(tau), _p(ssa), _p(pmom), _p(pe), _p(he), _p(te),
                                                       _p(albedo), _p(sza), _p(raa), _p(vza), _p(flux_factor), C.c_double(MISSING),
                                                       int(verbose), *[_p(o) for o in outs], C.byref(rc)),
          "ab200_vlidort_vector_lambert_multigeom")
    return (*outs, rc.value)


def vector_cx_multigeom(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, pe, he, te, U10m, V10m, mr, solar_zenith,
                        relat_azymuth, sensor_zenith, MISSING, verbose):
    """returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U (nobs,nch,ngeom), rc)"""
    ROT, tau, ssa, pmom, pe, he, te = (_F(x) for x in (ROT, tau, ssa, pmom, pe, he, te))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    sza, raa, vza, ngeom = _geom(solar_zenith, relat_azymuth, sensor_zenith, nobs)
    ch, dp, u, v, m = _v(channels), _v(depol), _v(U10m), _v(V10m), _v(mr)
    outs = [np.zeros((nobs, nch, ngeom), order="F") for _ in range(7)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_cx_multigeom(km, nch, nobs, ngeom, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol,
                                                  _p(ROT), _p(dp), _p(tau), _p(ssa), _p(pmom), _p(pe), _p(he), _p(te), _p(u), _p(v), _p(m),
                                                  _p(sza), _p(raa), _p(vza), C.c_double(MISSING), int(verbose), *[_p(o) for o in outs],
                                                  C.byref(rc)), "ab200_vlidort_vector_cx_multigeom")
    return (*outs, rc.value)


def vector_brdf_modis_multigeom(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, pe, he, te, kernel_wt, param,
                                solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose):
    """returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U (nobs,nch,ngeom), rc)"""
    ROT, tau, ssa, pmom, pe, he, te, kernel_wt, param = (_F(x) for x in (ROT, tau, ssa, pmom, pe, he, te, kernel_wt, param))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    sza, raa, vza, ngeom = _geom(solar_zenith, relat_azymuth, sensor_zenith, nobs)
    ch, dp = _v(channels), _v(depol)
    outs = [np.zeros((nobs, nch, ngeom), order="F") for _ in range(7)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_brdf_modis_multigeom(km, nch, nobs, ngeom, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom,
                                                          nPol, kernel_wt.shape[0], param.shape[0], _p(ROT), _p(dp), _p(tau), _p(ssa),
                                                          _p(pmom), _p(pe), _p(he), _p(te), _p(kernel_wt), _p(param), _p(sza), _p(raa),
                                                          _p(vza), C.c_double(MISSING), int(verbose), *[_p(o) for o in outs],
                                                          C.byref(rc)), "ab200_vlidort_vector_brdf_modis_multigeom")
    return (*outs, rc.value)


# ---- the remaining cloud forms of the reference's Fortran modules (no f2py wrapper exists for them; Fortran callers reach them through
#      the generic interfaces of fortran/ab200_rt_shims.F90).  Python names follow the f2py naming pattern of their siblings.
def scalar_lambert_cloud(channels, nstreams, plane_parallel, ROT, depol, alpha, tau, ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL,
                         pe, he, te, albedo, solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose):
    """VLIDORT_Scalar_Lambert_Cloud (GV/VLIDORT_LAMBERT.F90:524): clear-sky scalar runner with alpha and flux_factor; the cloud arrays do not
    enter (they may be None).  Returns (radiance, reflectance, rc)."""
    ROT, alpha, tau, ssa, pmom, pe, he, te, albedo, flux_factor = (_F(x) for x in (ROT, alpha, tau, ssa, pmom, pe, he, te, albedo, flux_factor))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(2)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_scalar_lambert_cloud(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, _p(ROT), _p(dp),
                                                   _p(alpha), _p(tau), _p(ssa), _opt(g), _p(pmom), _opt(tauI), _opt(ssaI), _opt(gI), _opt(pmomI),
                                                   _opt(tauL), _opt(ssaL), _opt(gL), _opt(pmomL), _p(pe), _p(he), _p(te), _p(albedo), _p(sza),
                                                   _p(raa), _p(vza), _p(flux_factor), C.c_double(MISSING), int(verbose), *[_p(o) for o in outs],
                                                   C.byref(rc)), "ab200_vlidort_scalar_lambert_cloud")
    return (*outs, rc.value)


def vector_lambert_cloud_multigeom(channels, nstreams, plane_parallel, ROT, depol, alpha, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL,
                                   pe, he, te, albedo, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose):
    """VLIDORT_Vector_Lambert_Cloud_MultiGeom (GV/VLIDORT_LAMBERT.F90:881): clear-sky multi-geometry runner with alpha, unit solar flux.
    Returns (radiance, reflectance, Q, U (nobs,nch,ngeom), rc)."""
    ROT, alpha, tau, ssa, pmom, pe, he, te, albedo = (_F(x) for x in (ROT, alpha, tau, ssa, pmom, pe, he, te, albedo))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    sza, raa, vza, ngeom = _geom(solar_zenith, relat_azymuth, sensor_zenith, nobs)
    ch, dp = _v(channels), _v(depol)
    outs = [np.zeros((nobs, nch, ngeom), order="F") for _ in range(4)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_lambert_cloud_multigeom(km, nch, nobs, ngeom, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol,
                                                             _p(ROT), _p(dp), _p(alpha), _p(tau), _p(ssa), _p(pmom), _opt(tauI), _opt(ssaI),
                                                             _opt(pmomI), _opt(tauL), _opt(ssaL), _opt(pmomL), _p(pe), _p(he), _p(te), _p(albedo),
                                                             _p(sza), _p(raa), _p(vza), C.c_double(MISSING), int(verbose),
                                                             *[_p(o) for o in outs], C.byref(rc)), "ab200_vlidort_vector_lambert_cloud_multigeom")
    return (*outs, rc.value)


def vector_cx_cloud_multigeom(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te,
                              U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose):
    """VLIDORT_Vector_CX_Cloud_MultiGeom (GV/VLIDORT_BRDF_CX.F90:1147): clear-sky runner, cloud arrays ignored as in the reference.
    Returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U (nobs,nch,ngeom), rc)."""
    ROT, tau, ssa, pmom, pe, he, te = (_F(x) for x in (ROT, tau, ssa, pmom, pe, he, te))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    sza, raa, vza, ngeom = _geom(solar_zenith, relat_azymuth, sensor_zenith, nobs)
    ch, dp, u, v, m = _v(channels), _v(depol), _v(U10m), _v(V10m), _v(mr)
    outs = [np.zeros((nobs, nch, ngeom), order="F") for _ in range(7)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_cx_cloud_multigeom(km, nch, nobs, ngeom, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol,
                                                        _p(ROT), _p(dp), _p(tau), _p(ssa), _p(pmom), _opt(tauI), _opt(ssaI), _opt(pmomI),
                                                        _opt(tauL), _opt(ssaL), _opt(pmomL), _p(pe), _p(he), _p(te), _p(u), _p(v), _p(m), _p(sza),
                                                        _p(raa), _p(vza), C.c_double(MISSING), int(verbose), *[_p(o) for o in outs],
                                                        C.byref(rc)), "ab200_vlidort_vector_cx_cloud_multigeom")
    return (*outs, rc.value)


def vector_brdf_modis_cloud_multigeom(channels, nstreams, plane_parallel, ROT, depol, alpha, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL,
                                      pmomL, pe, he, te, kernel_wt, param, solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING,
                                      verbose):
    """VLIDORT_Vector_LandMODIS_Cloud_MultiGeom (GV/VLIDORT_BRDF_MODIS.F90:1055): cloud runner for all geometries of an observation.
    Returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U (nobs,nch,ngeom), rc)."""
    ROT, alpha, tau, ssa, pmom, pe, he, te, tauI, ssaI, pmomI, tauL, ssaL, pmomL, kernel_wt, param, flux_factor = (
        _F(x) for x in (ROT, alpha, tau, ssa, pmom, pe, he, te, tauI, ssaI, pmomI, tauL, ssaL, pmomL, kernel_wt, param, flux_factor))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    assert tauI.shape == tau.shape and tauL.shape == tau.shape and pmomI.shape == pmom.shape and pmomL.shape == pmom.shape
    sza, raa, vza, ngeom = _geom(solar_zenith, relat_azymuth, sensor_zenith, nobs)
    ch, dp = _v(channels), _v(depol)
    outs = [np.zeros((nobs, nch, ngeom), order="F") for _ in range(7)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_brdf_modis_cloud_multigeom(km, nch, nobs, ngeom, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol,
                                                                kernel_wt.shape[0], param.shape[0], _p(ROT), _p(dp), _p(alpha), _p(tau), _p(ssa),
                                                                _p(pmom), _p(tauI), _p(ssaI), _p(pmomI), _p(tauL), _p(ssaL), _p(pmomL), _p(pe),
                                                                _p(he), _p(te), _p(kernel_wt), _p(param), _p(sza), _p(raa), _p(vza),
                                                                _p(flux_factor), C.c_double(MISSING), int(verbose), *[_p(o) for o in outs],
                                                                C.byref(rc)), "ab200_vlidort_vector_brdf_modis_cloud_multigeom")
    return (*outs, rc.value)


def vector_ocigisscx_cloud(channels, nstreams, plane_parallel, ROT, depol, alpha, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te,
                           U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose):
    """VLIDORT_Vector_OCIGissCX_Cloud (GV/VLIDORT_BRDF_CX.F90:2468): OCI GISS Cox-Munk ocean, trace gases, solar flux, aerosol + ice + liquid
    cloud layers.  Returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U, rc)."""
    ROT, alpha, tau, ssa, pmom, pe, he, te, tauI, ssaI, pmomI, tauL, ssaL, pmomL, flux_factor = (
        _F(x) for x in (ROT, alpha, tau, ssa, pmom, pe, he, te, tauI, ssaI, pmomI, tauL, ssaL, pmomL, flux_factor))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    assert tauI.shape == tau.shape and tauL.shape == tau.shape and pmomI.shape == pmom.shape and pmomL.shape == pmom.shape
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    u, v, m = _v(U10m), _v(V10m), _v(mr)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(7)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_ocigisscx_cloud(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, _p(ROT), _p(dp),
                                                     _p(alpha), _p(tau), _p(ssa), _p(pmom), _p(tauI), _p(ssaI), _p(pmomI), _p(tauL), _p(ssaL),
                                                     _p(pmomL), _p(pe), _p(he), _p(te), _p(u), _p(v), _p(m), _p(sza), _p(raa), _p(vza),
                                                     _p(flux_factor), C.c_double(MISSING), int(verbose), *[_p(o) for o in outs], C.byref(rc)),
          "ab200_vlidort_vector_ocigisscx_cloud")
    return (*outs, rc.value)


def scalar_ocigisscx_cloud(channels, nstreams, plane_parallel, ROT, depol, alpha, tau, ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL,
                           pe, he, te, U10m, V10m, mr, solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING, verbose):
    """VLIDORT_Scalar_OCIGissCX_Cloud (GV/VLIDORT_BRDF_CX.F90:2084): returns (radiance, reflectance, BRDF (nobs,nch), rc)."""
    arrs = [_F(x) for x in (ROT, alpha, tau, ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te)]
    km, nch, nobs, nMom, nPol = _dims(arrs[2], arrs[5])
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    u, v, m, ff = _v(U10m), _v(V10m), _v(mr), _F(flux_factor)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(3)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_scalar_ocigisscx_cloud(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, _p(arrs[0]),
                                                     _p(dp), *[_p(a) for a in arrs[1:]], _p(u), _p(v), _p(m), _p(sza), _p(raa), _p(vza), _p(ff),
                                                     C.c_double(MISSING), int(verbose), *[_p(o) for o in outs], C.byref(rc)),
          "ab200_vlidort_scalar_ocigisscx_cloud")
    return (*outs, rc.value)


def vector_brdf_modis_bpdf_cloud(channels, nstreams, plane_parallel, ROT, depol, alpha, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe,
                                 he, te, kernel_wt, RTLSparam, BPDFparam, solar_zenith, relat_azymuth, sensor_zenith, flux_factor, MISSING,
                                 verbose):
    """VLIDORT_Vector_LandMODIS_BPDF_Cloud (GV/VLIDORT_BRDF_MODIS_BPDF.F90:232): "place holder" in the reference -- clear-sky runner with alpha
    and flux_factor, the cloud arrays do not enter (they may be None).  Returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U, rc)."""
    ROT, alpha, tau, ssa, pmom, pe, he, te, kernel_wt, RTLSparam, BPDFparam, flux_factor = (
        _F(x) for x in (ROT, alpha, tau, ssa, pmom, pe, he, te, kernel_wt, RTLSparam, BPDFparam, flux_factor))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(7)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_brdf_modis_bpdf_cloud(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol,
                                                           kernel_wt.shape[0], RTLSparam.shape[0], BPDFparam.shape[0], _p(ROT), _p(dp), _p(alpha),
                                                           _p(tau), _p(ssa), _p(pmom), _opt(tauI), _opt(ssaI), _opt(pmomI), _opt(tauL), _opt(ssaL),
                                                           _opt(pmomL), _p(pe), _p(he), _p(te), _p(kernel_wt), _p(RTLSparam), _p(BPDFparam),
                                                           _p(sza), _p(raa), _p(vza), _p(flux_factor), C.c_double(MISSING), int(verbose),
                                                           *[_p(o) for o in outs], C.byref(rc)), "ab200_vlidort_vector_brdf_modis_bpdf_cloud")
    return (*outs, rc.value)


def vector_lambert_bpdf_cloud(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te,
                              albedo, BPDFparam, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose):
    """VLIDORT_Vector_Lambert_BPDF_Cloud (GV/VLIDORT_LAMBERT_BPDF.F90:228): clear-sky runner, cloud arrays ignored as in the reference.
    Returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U, rc)."""
    ROT, tau, ssa, pmom, pe, he, te, albedo, BPDFparam = (_F(x) for x in (ROT, tau, ssa, pmom, pe, he, te, albedo, BPDFparam))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(7)]
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_lambert_bpdf_cloud(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol,
                                                        BPDFparam.shape[0], _p(ROT), _p(dp), _p(tau), _p(ssa), _p(pmom), _opt(tauI), _opt(ssaI),
                                                        _opt(pmomI), _opt(tauL), _opt(ssaL), _opt(pmomL), _p(pe), _p(he), _p(te), _p(albedo),
                                                        _p(BPDFparam), _p(sza), _p(raa), _p(vza), C.c_double(MISSING), int(verbose),
                                                        *[_p(o) for o in outs], C.byref(rc)), "ab200_vlidort_vector_lambert_bpdf_cloud")
    return (*outs, rc.value)


def vector_ocigisscx_nobm_cloud(channels, nstreams, plane_parallel, ROT, depol, alpha, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he,
                                te, U10m, V10m, mr, sleave, sleave_iso, sleave_adjust, solar_zenith, relat_azymuth, sensor_zenith, flux_factor,
                                MISSING, verbose):
    """f2py VECTOR_OCIGissCX_NOBM_Cloud (VLIDORT_POLAR_py.F90:1055): OCI GISS Cox-Munk ocean + clouds + NOBM water-leaving radiance
    sleave (nch,nobs).  Returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U, rc, ADJUSTED_SLEAVE).  With sleave_adjust False VLIDORT 2.8.3
    drops the water-leaving term (its transmittance factor stays zero, see include/aeroapps_b200_rt.h) and so does this entry;
    sleave_adjust True (the transmittance iteration) is refused with a RuntimeError."""
    ROT, alpha, tau, ssa, pmom, pe, he, te, tauI, ssaI, pmomI, tauL, ssaL, pmomL, flux_factor, sleave = (
        _F(x) for x in (ROT, alpha, tau, ssa, pmom, pe, he, te, tauI, ssaI, pmomI, tauL, ssaL, pmomL, flux_factor, sleave))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    assert tauI.shape == tau.shape and tauL.shape == tau.shape and pmomI.shape == pmom.shape and pmomL.shape == pmom.shape
    assert sleave.shape == (nch, nobs), "sleave must be (nch,nobs)"
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    u, v, m = _v(U10m), _v(V10m), _v(mr)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(7)]
    adj = np.zeros((nobs, nch), order="F")
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_ocigisscx_nobm_cloud(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, _p(ROT),
                                                          _p(dp), _p(alpha), _p(tau), _p(ssa), _p(pmom), _p(tauI), _p(ssaI), _p(pmomI), _p(tauL),
                                                          _p(ssaL), _p(pmomL), _p(pe), _p(he), _p(te), _p(u), _p(v), _p(m), _p(sleave),
                                                          int(bool(sleave_iso)), int(bool(sleave_adjust)), _p(sza), _p(raa), _p(vza),
                                                          _p(flux_factor), C.c_double(MISSING), int(verbose), *[_p(o) for o in outs], C.byref(rc),
                                                          _p(adj)), "ab200_vlidort_vector_ocigisscx_nobm_cloud")
    return (*outs, rc.value, adj)


def vector_gisscx_nobm_cloud(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, tauI, ssaI, pmomI, tauL, ssaL, pmomL, pe, he, te, U10m,
                             V10m, mr, sleave, sleave_iso, sleave_adjust, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose):
    """VLIDORT_Vector_GissCX_NOBM_Cloud (GV/VLIDORT_BRDF_CX.F90:3102): returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U, rc,
    ADJUSTED_SLEAVE); sleave_adjust as for vector_ocigisscx_nobm_cloud."""
    ROT, tau, ssa, pmom, pe, he, te, tauI, ssaI, pmomI, tauL, ssaL, pmomL, sleave = (
        _F(x) for x in (ROT, tau, ssa, pmom, pe, he, te, tauI, ssaI, pmomI, tauL, ssaL, pmomL, sleave))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    assert sleave.shape == (nch, nobs), "sleave must be (nch,nobs)"
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    u, v, m = _v(U10m), _v(V10m), _v(mr)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(7)]
    adj = np.zeros((nobs, nch), order="F")
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_gisscx_nobm_cloud(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, _p(ROT), _p(dp),
                                                       _p(tau), _p(ssa), _p(pmom), _p(tauI), _p(ssaI), _p(pmomI), _p(tauL), _p(ssaL), _p(pmomL),
                                                       _p(pe), _p(he), _p(te), _p(u), _p(v), _p(m), _p(sleave), int(bool(sleave_iso)),
                                                       int(bool(sleave_adjust)), _p(sza), _p(raa), _p(vza), C.c_double(MISSING), int(verbose),
                                                       *[_p(o) for o in outs], C.byref(rc), _p(adj)), "ab200_vlidort_vector_gisscx_nobm_cloud")
    return (*outs, rc.value, adj)


def vector_cx_nobm(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, pe, he, te, U10m, V10m, mr, sleave, sleave_iso, sleave_adjust,
                   solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose):
    """VLIDORT_Vector_CX_NOBM (GV/VLIDORT_BRDF_CX.F90:1351): returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U, rc, ADJUSTED_SLEAVE)."""
    ROT, tau, ssa, pmom, pe, he, te, sleave = (_F(x) for x in (ROT, tau, ssa, pmom, pe, he, te, sleave))
    km, nch, nobs, nMom, nPol = _dims(tau, pmom)
    assert sleave.shape == (nch, nobs), "sleave must be (nch,nobs)"
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    u, v, m = _v(U10m), _v(V10m), _v(mr)
    outs = [np.zeros((nobs, nch), order="F") for _ in range(7)]
    adj = np.zeros((nobs, nch), order="F")
    rc = C.c_int(0)
    check(lib().ab200_vlidort_vector_cx_nobm(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, _p(ROT), _p(dp), _p(tau),
                                             _p(ssa), _p(pmom), _p(pe), _p(he), _p(te), _p(u), _p(v), _p(m), _p(sleave), int(bool(sleave_iso)),
                                             int(bool(sleave_adjust)), _p(sza), _p(raa), _p(vza), C.c_double(MISSING), int(verbose),
                                             *[_p(o) for o in outs], C.byref(rc), _p(adj)), "ab200_vlidort_vector_cx_nobm")
    return (*outs, rc.value, adj)


def scalar_gisscx_nobm_cloud(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he,
                             te, U10m, V10m, mr, sleave, sleave_iso, sleave_adjust, solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose):
    """VLIDORT_Scalar_GissCX_NOBM_Cloud (GV/VLIDORT_BRDF_CX.F90:2671): returns (radiance, reflectance, BRDF, rc, ADJUSTED_SLEAVE)."""
    arrs = [_F(x) for x in (ROT, tau, ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te)]
    km, nch, nobs, nMom, nPol = _dims(arrs[1], arrs[4])
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    u, v, m, sl = _v(U10m), _v(V10m), _v(mr), _F(sleave)
    assert sl.shape == (nch, nobs), "sleave must be (nch,nobs)"
    outs = [np.zeros((nobs, nch), order="F") for _ in range(3)]
    adj = np.zeros((nobs, nch), order="F")
    rc = C.c_int(0)
    check(lib().ab200_vlidort_scalar_gisscx_nobm_cloud(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, _p(arrs[0]),
                                                       _p(dp), *[_p(a) for a in arrs[1:]], _p(u), _p(v), _p(m), _p(sl), int(bool(sleave_iso)),
                                                       int(bool(sleave_adjust)), _p(sza), _p(raa), _p(vza), C.c_double(MISSING), int(verbose),
                                                       *[_p(o) for o in outs], C.byref(rc), _p(adj)), "ab200_vlidort_scalar_gisscx_nobm_cloud")
    return (*outs, rc.value, adj)


def scalar_ocigisscx_nobm_cloud(channels, nstreams, plane_parallel, ROT, depol, alpha, tau, ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL,
                                pmomL, pe, he, te, U10m, V10m, mr, sleave, sleave_iso, sleave_adjust, solar_zenith, relat_azymuth, sensor_zenith,
                                flux_factor, MISSING, verbose):
    """VLIDORT_Scalar_OCIGissCX_NOBM_Cloud (GV/VLIDORT_BRDF_CX.F90:2883): returns (radiance, reflectance, BRDF, rc, ADJUSTED_SLEAVE)."""
    arrs = [_F(x) for x in (ROT, alpha, tau, ssa, g, pmom, tauI, ssaI, gI, pmomI, tauL, ssaL, gL, pmomL, pe, he, te)]
    km, nch, nobs, nMom, nPol = _dims(arrs[2], arrs[5])
    ch, dp, sza, raa, vza = _v(channels), _v(depol), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    u, v, m, sl, ff = _v(U10m), _v(V10m), _v(mr), _F(sleave), _F(flux_factor)
    assert sl.shape == (nch, nobs), "sleave must be (nch,nobs)"
    outs = [np.zeros((nobs, nch), order="F") for _ in range(3)]
    adj = np.zeros((nobs, nch), order="F")
    rc = C.c_int(0)
    check(lib().ab200_vlidort_scalar_ocigisscx_nobm_cloud(km, nch, nobs, _p(ch), int(nstreams), int(bool(plane_parallel)), nMom, nPol, _p(arrs[0]),
                                                          _p(dp), *[_p(a) for a in arrs[1:]], _p(u), _p(v), _p(m), _p(sl), int(bool(sleave_iso)),
                                                          int(bool(sleave_adjust)), _p(sza), _p(raa), _p(vza), _p(ff), C.c_double(MISSING),
                                                          int(verbose), *[_p(o) for o in outs], C.byref(rc), _p(adj)),
          "ab200_vlidort_scalar_ocigisscx_nobm_cloud")
    return (*outs, rc.value, adj)


def last_diag():
    """(nfourier, status) per pixel p = j*nch + i (per pixel-view (j*nch + i)*ngeom + g for the multi-geometry forms) of this
    thread's most recent VLIDORT call -- diagnostics, not in the reference."""
    lib().ab200_vlidort_last_diag.restype = C.c_long
    n = lib().ab200_vlidort_last_diag(None, None, C.c_long(0))
    nf = np.zeros(n, np.int32); st = np.zeros(n, np.int32)
    lib().ab200_vlidort_last_diag(nf.ctypes.data_as(C.POINTER(C.c_int)), st.ctypes.data_as(C.POINTER(C.c_int)), C.c_long(n))
    return nf, st


def stats(reset=False):
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    d, e, f = C.c_long(), C.c_long(), C.c_long()
    lib().ab200_vlidort_stats(C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e), C.byref(f), int(reset))
    return dict(ms_setup=a.value, ms_eigen=b.value, ms_bvp=c.value, n_eigen=d.value, n_bvp=e.value, npix=f.value)
