"""TEST INFRASTRUCTURE ONLY -- ctypes access to the wrapper-level oracle entry points (orc_vlidort_run,
orc_rot_calc) and a synthetic scene builder used by tests, smoke() and the bench CPU arms."""
import ctypes as C
import numpy as np

from . import vlidort_oracle as vo
from . import mie_oracle as mo

_dp = C.POINTER(C.c_double)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def F(a):
    return np.asfortranarray(a, dtype=np.float64)


def rot_calc(channels, pe, he, te):
    pe, he, te = F(pe), F(he), F(te)
    km = pe.shape[0] - 1; nobs = pe.shape[1]
    ch = np.ascontiguousarray(channels, dtype=np.float64); nch = len(ch)
    ROT = np.zeros((km, nobs, nch), order="F"); depol = np.zeros(nch)
    vo.lib().orc_rot_calc(km, nch, nobs, _p(ch), _p(pe), _p(he), _p(te), _p(ROT), _p(depol))
    return ROT, depol


def vlidort_run(nstokes, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, he, sza, raa, vza, albedo=None,
                alpha=None, flux_factor=None, MISSING=1e15, scalar_glue=False, brdf=None, skip=None, obs_range=None):
    """Wrapper-level oracle: returns dict(radiance, reflectance, Q, U (nobs,nch), nfourier, status (nobs,nch))."""
    ROT, tau, ssa, pmom, he = F(ROT), F(tau), F(ssa), F(pmom), F(he)
    km, nch, nobs = tau.shape
    nMom, nPol = pmom.shape[3], pmom.shape[4]
    depol = np.ascontiguousarray(depol, dtype=np.float64)
    sza, raa, vza = (np.ascontiguousarray(x, dtype=np.float64) for x in (sza, raa, vza))
    alb = F(albedo) if albedo is not None else None
    al = F(alpha) if alpha is not None else None
    ff = F(flux_factor) if flux_factor is not None else None
    out = {k: np.zeros((nobs, nch), order="F") for k in ("radiance", "reflectance", "Q", "U")}
    nf = np.zeros((nobs, nch), np.int32); st = np.zeros((nobs, nch), np.int32)
    b = {k: None for k in ("brdf_f", "brdf_f_0", "user_brdf_f", "user_brdf_f_0", "exactdb")}
    if brdf is not None:
        b = {k: np.ascontiguousarray(brdf[k], dtype=np.float64) for k in b}
    sk = None if skip is None else np.ascontiguousarray(skip, dtype=np.uint8)
    o0, o1 = obs_range if obs_range else (0, nobs)
    vo.lib().orc_vlidort_run(int(nstokes), int(scalar_glue), km, nch, nobs, int(nstreams), int(plane_parallel), nMom, nPol,
                             _p(ROT), _p(depol), _p(al), _p(tau), _p(ssa), _p(pmom), _p(he), _p(alb), _p(sza), _p(raa),
                             _p(vza), _p(ff), C.c_double(MISSING), 0 if brdf is None else 1, _p(b["brdf_f"]),
                             _p(b["brdf_f_0"]), _p(b["user_brdf_f"]), _p(b["user_brdf_f_0"]), _p(b["exactdb"]),
                             None if sk is None else sk.ctypes.data_as(C.POINTER(C.c_ubyte)), int(o0), int(o1),
                             _p(out["radiance"]), _p(out["reflectance"]), _p(out["Q"]), _p(out["U"]),
                             nf.ctypes.data_as(C.POINTER(C.c_int)), st.ctypes.data_as(C.POINTER(C.c_int)))
    out["nfourier"] = nf; out["status"] = st
    return out


def set_asymtx_tolerance(tol=1e-12):
    """ASYMTX tolerance of the wrapper-level oracle runs (reference value 1e-12)."""
    vo.lib().orc_set_asymtx_tolerance(C.c_double(tol))


_cloud_keep = None


def set_cloud(sc=None):
    """register (sc: dict with tauI, ssaI, pmomI, tauL, ssaL, pmomL, channels) or clear (None) the cloud inputs mixed into the glue of the
    following oracle runs: VLIDORT_Run_Vector_Cloud (V283/VLIDORT_ScatMod.F90:1865-2220)"""
    global _cloud_keep
    if sc is None:
        _cloud_keep = None
        vo.lib().orc_set_cloud(None, None, None, None, None, None, None)
        return
    keep = [F(sc[k]) for k in ("tauI", "ssaI", "pmomI", "tauL", "ssaL", "pmomL")] + [np.ascontiguousarray(sc["channels"], dtype=np.float64)]
    _cloud_keep = keep                                            # the oracle holds raw pointers: keep the arrays alive
    vo.lib().orc_set_cloud(*[_p(a) for a in keep])


class _OrcSurface(C.Structure):
    _fields_ = [("kind", C.c_int)] + [(n, C.POINTER(C.c_double)) for n in ("u10m", "v10m", "mr", "kernel_wt", "param", "BR", "BR_Q", "BR_U", "bpdf")]


SURFACE_KINDS = {"cx": 2, "gisscx": 3, "modis": 4, "ocicx": 5, "ocigisscx": 6, "bpdf": 7, "modis_bpdf": 8, "lambert_bpdf": 9}


def vlidort_run_surface(kind, nstokes, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, he, sza, raa, vza, u10m=None,
                        v10m=None, mr=None, kernel_wt=None, param=None, MISSING=1e15, scalar_glue=False, obs_range=None,
                        alpha=None, flux_factor=None, bpdf=None, albedo=None):
    """BRDF wrappers (VLIDORT_Vector_CX / _GissCX / _LandMODIS and their scalar forms): dict(radiance, reflectance, Q, U,
    BR, BR_Q, BR_U (nobs,nch), nfourier, status)."""
    ROT, tau, ssa, pmom, he = F(ROT), F(tau), F(ssa), F(pmom), F(he)
    km, nch, nobs = tau.shape
    nMom, nPol = pmom.shape[3], pmom.shape[4]
    depol = np.ascontiguousarray(depol, dtype=np.float64)
    sza, raa, vza = (np.ascontiguousarray(x, dtype=np.float64) for x in (sza, raa, vza))
    out = {k: np.zeros((nobs, nch), order="F") for k in ("radiance", "reflectance", "Q", "U", "BR", "BR_Q", "BR_U")}
    nf = np.zeros((nobs, nch), np.int32); st = np.zeros((nobs, nch), np.int32)
    keep = [np.ascontiguousarray(x, dtype=np.float64) if x is not None else None for x in (u10m, v10m, mr)]
    keep += [F(kernel_wt) if kernel_wt is not None else None, F(param) if param is not None else None]
    alpha = F(alpha) if alpha is not None else None
    flux_factor = F(flux_factor) if flux_factor is not None else None
    bpdf = F(bpdf) if bpdf is not None else None
    albedo = F(albedo) if (albedo is not None and kind == "lambert_bpdf") else None
    sf = _OrcSurface(SURFACE_KINDS[kind], *[_p(k) for k in keep], _p(out["BR"]), _p(out["BR_Q"]), _p(out["BR_U"]), _p(bpdf))
    o0, o1 = obs_range if obs_range else (0, nobs)
    vo.lib().orc_vlidort_run_surf(int(nstokes), int(scalar_glue), km, nch, nobs, int(nstreams), int(plane_parallel), nMom, nPol,
                                  _p(ROT), _p(depol), _p(alpha), _p(tau), _p(ssa), _p(pmom), _p(he), _p(albedo), _p(sza), _p(raa), _p(vza),
                                  _p(flux_factor), C.c_double(MISSING), 0, None, None, None, None, None, None, int(o0), int(o1),
                                  _p(out["radiance"]), _p(out["reflectance"]), _p(out["Q"]), _p(out["U"]),
                                  nf.ctypes.data_as(C.POINTER(C.c_int)), st.ctypes.data_as(C.POINTER(C.c_int)), C.byref(sf))
    out["nfourier"] = nf; out["status"] = st
    return out


def synthetic_scene(nobs, nch_nm=(550.0,), nmom=64, km=72, seed=7, tables=None, mt=None, sza_max=75.0, vza_max=65.0):
    """SURVEY.md section 8(d) synthetic pixels: GOCART-like columns -> Mie oracle (vector AOPs) -> edge variables ->
    Rayleigh; returns float64 Fortran arrays in the drop-in layouts."""
    if mt is None:
        from aeroapps_b200 import mietables as mt
    tables = tables or mt.synthetic_tables(nmom=nmom)
    st = mt.synthetic_state(nobs, km, seed=seed)
    chan_m = [c * 1e-9 for c in nch_nm]
    om = mo.OracleMie(chan_m, nmom, tables)
    tau, ssa, g, pmom, rc = om.getaopvector(list(nch_nm), st["vnames"], st["qm"], st["rh"], nmom, 6)
    assert rc == 0
    pe, ze, te = mo.getedgevars(st["airdens"], st["delp"], 1.0)
    pmom = pmom.astype(np.float64)
    pmom[..., 1] *= -1.0; pmom[..., 3] *= -1.0           # caller-side sign flip of P12, P34 (LV/py_leo_vlidort/vlidort.py:438-440)
    ROT, depol = rot_calc(nch_nm, pe, ze, te)
    rng = np.random.default_rng(seed + 1000)
    scene = dict(km=km, nobs=nobs, nch=len(nch_nm), channels=np.array(nch_nm, dtype=np.float64), nMom=nmom, nPol=6,
                 tau=F(tau), ssa=F(ssa), g=F(g), pmom=F(pmom), pe=F(pe), he=F(ze), te=F(te), ROT=F(ROT), depol=depol,
                 sza=rng.uniform(0.0, sza_max, nobs), vza=rng.uniform(0.0, vza_max, nobs), raa=rng.uniform(0.0, 180.0, nobs),
                 albedo=F(rng.uniform(0.02, 0.4, (nobs, len(nch_nm)))),
                 u10m=rng.normal(0, 5, nobs), v10m=rng.normal(0, 5, nobs), mr=np.full(len(nch_nm), 1.334),
                 kernel_wt=F(np.stack([rng.uniform(0.02, 0.3, (len(nch_nm), nobs)), rng.uniform(0.0, 0.05, (len(nch_nm), nobs)),
                                       rng.uniform(0.0, 0.15, (len(nch_nm), nobs))])),
                 param=F(np.stack([np.full((len(nch_nm), nobs), 2.0), np.full((len(nch_nm), nobs), 1.0)])))
    return scene


def vbrdf(nstokes, nstreams, which, factors, pars, lambertian, nstreams_brdf, do_shadow, sza, vza, raa):
    """VBRDF supplement for one geometry (oracle/vbrdf_oracle.cpp).  Returns the Fourier arrays in the layout
    vl_master consumes: brdf_f[m][q][i][j], brdf_f_0[m][q][i], user_brdf_f[m][q][j], user_brdf_f_0[m][q], exactdb[q]."""
    L = vo.lib()
    N, nk = int(nstreams), len(which)
    w = np.asarray(which, np.int32); f = np.asarray(factors, np.float64)
    p = np.ascontiguousarray(np.asarray(pars, np.float64).reshape(nk, 4)); lf = np.asarray(lambertian, np.int32)
    bf = np.zeros((2 * N, 16, N, N)); bf0 = np.zeros((2 * N, 16, N)); ubf = np.zeros((2 * N, 16, N)); ubf0 = np.zeros((2 * N, 16))
    ex = np.zeros(16); wsa = C.c_double(0.0)
    dp = C.POINTER(C.c_double); ip = C.POINTER(C.c_int)
    L.orc_vbrdf.restype = C.c_int
    st = L.orc_vbrdf(int(nstokes), N, nk, w.ctypes.data_as(ip), f.ctypes.data_as(dp), p.ctypes.data_as(dp), lf.ctypes.data_as(ip),
                     int(nstreams_brdf), int(bool(do_shadow)), C.c_double(sza), C.c_double(vza), C.c_double(raa),
                     bf.ctypes.data_as(dp), bf0.ctypes.data_as(dp), ubf.ctypes.data_as(dp), ubf0.ctypes.data_as(dp),
                     ex.ctypes.data_as(dp), C.byref(wsa))
    return dict(status=st, brdf_f=bf, brdf_f_0=bf0, user_brdf_f=ubf, user_brdf_f_0=ubf0, exactdb=ex, wsa=wsa.value)
