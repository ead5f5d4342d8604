"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/libangles_oracle.so (CPU restatement of GeoAngles_ / LidarAngles_ / stnAngles_).
Imported by tests/ only; never by the product package aeroapps_b200/.  PARITY UNPINNED (see angles_oracle.c header)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
_ip, _dp = C.POINTER(C.c_int), C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", _HERE, "libangles_oracle.so"])
        _lib = C.CDLL(os.path.join(_HERE, "libangles_oracle.so"))
    return _lib


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def angles(mode, yr, mo, day, hr, mn, sec, latp, lonp, latss=None, lonss=None, latss0=0.0, lonss0=0.0, tz=0.0, hgtss=35786.0):
    """mode 0: satangles, 1: satangles_fromspace, 2: solar angles only -> (vaa, vza, saa, sza) or (saa, sza)"""
    yr, mo, day, hr, mn = (_i(x) for x in (yr, mo, day, hr, mn))
    sec, latp, lonp = _d(sec), _d(latp), _d(lonp)
    n = len(yr)
    la = None if latss is None else _d(latss)
    lo = None if lonss is None else _d(lonss)
    out = [np.zeros(n) for _ in range(4)]
    P = lambda a, t: None if a is None else a.ctypes.data_as(t)
    lib().orc_angles(n, int(mode), P(yr, _ip), P(mo, _ip), P(day, _ip), P(hr, _ip), P(mn, _ip), P(sec, _dp), P(latp, _dp), P(lonp, _dp),
                     P(la, _dp), P(lo, _dp), C.c_double(latss0), C.c_double(lonss0), C.c_double(tz), C.c_double(hgtss),
                     *[P(o, _dp) for o in out])
    return tuple(out[2:]) if mode == 2 else tuple(out)
