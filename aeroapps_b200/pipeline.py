"""Chained column pipeline: the call sequence of leo_vlidort / polar_vlidort
    pe, ze, te = MieObs_.getEdgeVars(airdens, delp, ptop)
    ROT, depol = VLIDORT_POLAR_.ROT_CALC(channels, pe, ze, te, ...)
    tau, ssa, g, pmom = MieObs_.getAOPvector(...);  pmom[..., 1] *= -1;  pmom[..., 3] *= -1
    I, refl, BR, Q, U, BR_Q, BR_U = VLIDORT_POLAR_.VECTOR_CX(channels, nstreams, plane_parallel, ROT, depol, tau, ssa, pmom, ...)
(LV/polar_vlidort.py:600-950, LV/py_leo_vlidort/vlidort.py:400-580) as ONE C-ABI call, `ab200_pipeline_vector_cx`: only the
model state and the geometry cross PCIe, every intermediate stays in HBM.  Bit-identical to the four separate calls."""
import ctypes as C

import numpy as np

from ._lib import lib, check
from .MieObs_ import _f32, _names, _p32


def _v(a):
    return np.ascontiguousarray(a, dtype=np.float64).reshape(-1)


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def vector_cx_pipeline(mie, channels, vname, qm, rh, airdens, delp, ptop, nstreams, plane_parallel, nmom, npol, U10m, V10m, mr,
                       solar_zenith, relat_azymuth, sensor_zenith, MISSING, verbose):
    """mie: MieObs_.ChemMie; channels [nm]; qm (km,nq,nobs), rh / airdens / delp (km,nobs) float32; U10m, V10m, angles (nobs),
    mr (nch).  Returns (radiance, reflectance, BR, Q, U, BR_Q, BR_U, rc), each (nobs,nch) float64."""
    qm, rh, airdens, delp = _f32(qm), _f32(rh), _f32(airdens), _f32(delp)
    km, nq, nobs = qm.shape
    ch = _v(channels)
    nch = len(ch)
    vn, nqn = _names(vname)
    assert nqn == nq and rh.shape == (km, nobs) and airdens.shape == (km, nobs) and delp.shape == (km, nobs)
    u, v, m, sza, raa, vza = _v(U10m), _v(V10m), _v(mr), _v(solar_zenith), _v(relat_azymuth), _v(sensor_zenith)
    assert len(u) == nobs and len(v) == nobs and len(m) == nch and len(sza) == nobs and len(raa) == nobs and len(vza) == nobs
    outs = [np.zeros((nobs, nch), order="F") for _ in range(7)]
    rc = C.c_int(0)
    check(lib().ab200_pipeline_vector_cx(mie.handle, km, nobs, nch, nq, _p(ch), vn.ctypes.data_as(C.c_char_p), _p32(qm), _p32(rh),
                                         _p32(airdens), _p32(delp), C.c_float(ptop), int(nstreams), int(bool(plane_parallel)),
                                         int(nmom), int(npol), _p(u), _p(v), _p(m), _p(sza), _p(raa), _p(vza), C.c_double(MISSING),
                                         int(verbose), *[_p(o) for o in outs], C.byref(rc)), "ab200_pipeline_vector_cx")
    return (*outs, rc.value)
