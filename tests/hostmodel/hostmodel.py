"""TEST-ONLY: builds and calls tests/hostmodel/libvl_hostmodel.so (the product's kernel bodies compiled for the CPU
with a one-lane team) so that their mathematics can be compared with the oracle without a GPU."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvl_hostmodel.so")
_dp = C.POINTER(C.c_double)
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "vl_hostmodel.cpp")
        hdrs = [os.path.join(_HERE, "..", "..", "aeroapps_b200", "csrc", f) for f in ("vl_common.cuh", "vl_setup.cuh", "vl_eigen.cuh", "vl_bvp.cuh", "vl_brdf.cuh", "vl_linalg.cuh", "vl_layerops.cuh", "vl_add.cuh")]
        newest = max(os.path.getmtime(f) for f in [src] + hdrs)
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", _SO, src])
        _lib = C.CDLL(_SO)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def run(nstokes, nstreams, plane_parallel, sc, lambertian=True, scalar_glue=False, nfine=3, nadir_from_obs=None,
        accuracy=float(np.float32(0.0001)), brdf=None, alpha=None, flux_factor=None, max_fourier=-1, surface=None, ngeom=1):
    """ngeom > 1: sc["raa"], sc["vza"] are (nobs, ngeom); results get a trailing ngeom axis"""
    F = lambda a: np.asfortranarray(a, dtype=np.float64)
    km, nch, nobs = sc["tau"].shape
    npix = nobs * nch * ngeom
    stokes = np.zeros((npix, 4)); nf = np.zeros(npix, np.int32); st = np.zeros(npix, np.int32)
    b = {k: None for k in ("brdf_f", "brdf_f_0", "user_brdf_f", "user_brdf_f_0", "exactdb")}
    if brdf is not None:
        b = {k: np.ascontiguousarray(brdf[k], dtype=np.float64) for k in b}
    arrs = [F(sc["ROT"]), np.ascontiguousarray(sc["depol"], dtype=np.float64), None if alpha is None else F(alpha), F(sc["tau"]),
            F(sc["ssa"]), F(sc["pmom"]), F(sc["he"]), F(sc["albedo"])] + [np.ascontiguousarray(np.asarray(sc[k], dtype=np.float64).ravel(order="F")) for k in ("sza", "raa", "vza")]
    ff = None if flux_factor is None else F(flux_factor)
    kinds = {None: 0, "cx": 2, "gisscx": 3, "modis": 4, "ocicx": 5, "ocigisscx": 6}
    sk = kinds[surface]
    sarr = [None] * 5
    if sk in (2, 3, 5, 6):
        sarr[:3] = [np.ascontiguousarray(sc[k], dtype=np.float64) for k in ("u10m", "v10m", "mr")]
    elif sk == 4:
        sarr[3:] = [F(sc["kernel_wt"]), F(sc["param"])]
    BR = np.zeros(npix); wsa = np.zeros(npix, np.int32)
    lib().hm_vlidort(int(ngeom), int(nstokes), int(nstreams), km, nch, nobs, sc["nMom"], sc["nPol"], int(plane_parallel), int(lambertian),
                     int(scalar_glue), int(nfine), nobs if nadir_from_obs is None else int(nadir_from_obs), C.c_double(6371.0),
                     C.c_double(accuracy), *[_p(a) for a in arrs], _p(ff), _p(b["brdf_f"]), _p(b["brdf_f_0"]), _p(b["user_brdf_f"]),
                     _p(b["user_brdf_f_0"]), _p(b["exactdb"]), int(max_fourier), sk, *[_p(a) for a in sarr], _p(BR),
                     wsa.ctypes.data_as(C.POINTER(C.c_int)), _p(stokes), nf.ctypes.data_as(C.POINTER(C.c_int)),
                     st.ctypes.data_as(C.POINTER(C.c_int)))
    if ngeom > 1:
        out = (stokes.reshape(nobs, nch, ngeom, 4), nf.reshape(nobs, nch, ngeom), st.reshape(nobs, nch, ngeom))
        return out + (BR.reshape(nobs, nch, ngeom),) if surface is not None else out
    if surface is not None:
        return stokes.reshape(nobs, nch, 4), nf.reshape(nobs, nch), st.reshape(nobs, nch), BR.reshape(nobs, nch), wsa[:nobs * nch].reshape(nobs, nch)
    return stokes.reshape(nobs, nch, 4), nf.reshape(nobs, nch), st.reshape(nobs, nch)
