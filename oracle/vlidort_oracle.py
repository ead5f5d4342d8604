"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/libvlidort_oracle.so (CPU restatement of the
VLIDORT 2.8.3 path used by AeroApps).  Never imported by the product package."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvlidort_oracle.so")
_dp = C.POINTER(C.c_double)


class VLInput(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("nstokes", "nstreams", "nlayers", "nmoments_input", "nfine",
                                       "do_plane_parallel", "do_focorr", "do_focorr_nadir", "do_focorr_outgoing",
                                       "do_deltam", "do_lambertian", "max_fourier_override")] + \
               [(n, C.c_double) for n in ("accuracy", "asymtx_tolerance", "flux_factor", "earth_radius", "albedo",
                                          "sza", "vza", "raa")] + \
               [(n, _dp) for n in ("height_grid", "deltau", "omega", "greekmat", "brdf_f", "brdf_f_0", "user_brdf_f",
                                   "user_brdf_f_0", "exactdb")]


class VLOutput(C.Structure):
    _fields_ = [("stokes", C.c_double * 4), ("stokes_ss", C.c_double * 4), ("stokes_db", C.c_double * 4),
                ("nfourier", C.c_int), ("status", C.c_int)]


_lib = None
_native = None
_use_native = False


def _cpu_tag():
    """identifies the host CPU: a -march=native build must never run on another machine (the tree travels to the GPU box)"""
    import hashlib
    try:
        txt = open("/proc/cpuinfo").read()
        key = "".join(l for l in txt.splitlines() if l.startswith(("model name", "flags")))[:20000]
    except OSError:
        key = "unknown"
    return hashlib.sha1(key.encode()).hexdigest()[:12]


def use_native(on=True):
    """bench.py's CPU arms: time a -O3 -march=native build of the same sources (built on first use, on the host that runs it, into
    oracle/_native/).  The portable -O2 build stays the checker of the parity tests.  Returns True when the native build is active."""
    global _native, _use_native
    if on and _native is None:
        d = os.path.join(_HERE, "_native")
        so = os.path.join(d, f"libvlidort_oracle_{_cpu_tag()}.so")
        try:
            if not os.path.exists(so):
                os.makedirs(d, exist_ok=True)
                subprocess.check_call(["g++", "-O3", "-march=native", "-std=c++17", "-fPIC", "-shared", "-o", so,
                                       os.path.join(_HERE, "vlidort_oracle.cpp"), os.path.join(_HERE, "vbrdf_oracle.cpp"), "-lm"])
            _native = C.CDLL(so)
        except Exception:
            _native = False
    _use_native = bool(on and _native)
    return _use_native


def lib():
    global _lib
    if _use_native:
        return _native
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-s", "-C", _HERE])
        _lib = C.CDLL(_SO)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def vlidort_master(nstokes, nstreams, height_grid, deltau, omega, greekmat, sza, vza, raa, albedo=0.0,
                   lambertian=True, plane_parallel=False, focorr=True, focorr_nadir=False, deltam=True, nfine=3,
                   accuracy=float(np.float32(0.0001)), flux_factor=1.0, earth_radius=6371.0, max_fourier=-1,
                   brdf=None, asymtx_tolerance=1e-12):
    """greekmat: (nmom+1, nlayers, 16) C-contiguous double.  brdf: dict with brdf_f (2N,16,N,N), brdf_f_0 (2N,16,N),
    user_brdf_f (2N,16,N), user_brdf_f_0 (2N,16), exactdb (16).  Returns (stokes[4], nfourier, status, ss, db)."""
    g = np.ascontiguousarray(greekmat, dtype=np.float64)
    h = np.ascontiguousarray(height_grid, dtype=np.float64)
    d = np.ascontiguousarray(deltau, dtype=np.float64)
    w = np.ascontiguousarray(omega, dtype=np.float64)
    nl = len(d)
    assert g.shape[1] == nl and g.shape[2] == 16 and len(h) == nl + 1
    inp = VLInput(nstokes, nstreams, nl, g.shape[0] - 1, nfine, int(plane_parallel), int(focorr), int(focorr_nadir),
                  int(focorr and not focorr_nadir), int(deltam), int(lambertian), int(max_fourier), accuracy,
                  asymtx_tolerance, flux_factor, earth_radius, albedo, sza, vza, raa, _p(h), _p(d), _p(w), _p(g))
    keep = []
    if not lambertian:
        for k in ("brdf_f", "brdf_f_0", "user_brdf_f", "user_brdf_f_0", "exactdb"):
            a = np.ascontiguousarray(brdf[k], dtype=np.float64); keep.append(a); setattr(inp, k, _p(a))
    out = VLOutput()
    lib().vl_master(C.byref(inp), C.byref(out))
    return (np.array(out.stokes[:]), out.nfourier, out.status, np.array(out.stokes_ss[:]), np.array(out.stokes_db[:]))
