"""ctypes mirrors of the reference's three geometry f2py modules, same function names and argument order:
    GeoAngles_.satangles(yr, mo, day, hr, min, sec, latp, lonp, lonss)                    rtms/geo_vlidort/GeoAngles_py.F90:682
    LidarAngles_.satangles(yr, mo, day, hr, min, sec, latp, lonp, latss, lonss, tz, hgtss)  rtms/leo_vlidort/LidarAngles_py.F90:901
    LidarAngles_.satangles_fromspace(...same...)                                           :956
    LidarAngles_.solarangles(yr, mo, day, hr, min, sec, latp, lonp, tz)                     :1011
    stnAngles_.sunangles(yr, mo, day, hr, min, sec, latp, lonp, tz)                         rtms/stn_vlidort/stnAngles_py.F90:338
A driver switches with  `from aeroapps_b200.angles import GeoAngles_`  (or LidarAngles_, stnAngles_).  The arithmetic runs on the GPU
(csrc/geo_angles.cu); there is no CPU fallback."""
import ctypes as C
import types

import numpy as np

from ._lib import lib, check

_ip, _dp = C.POINTER(C.c_int), C.POINTER(C.c_double)


def _i(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.int32).reshape(-1)
    assert n is None or len(a) == n, "angle calculators: array lengths differ"
    return a


def _d(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    assert n is None or len(a) == n, "angle calculators: array lengths differ"
    return a


def _common(yr, mo, day, hr, mn, sec, latp, lonp):
    yr = _i(yr); n = len(yr)
    return n, [yr, _i(mo, n), _i(day, n), _i(hr, n), _i(mn, n)], [_d(sec, n), _d(latp, n), _d(lonp, n)]


def _call(fn, name, n, ints, dbls, extra, nout):
    outs = [np.zeros(n) for _ in range(nout)]
    check(fn(n, *[a.ctypes.data_as(_ip) for a in ints], *[a.ctypes.data_as(_dp) for a in dbls], *extra,
             *[o.ctypes.data_as(_dp) for o in outs]), name)
    return tuple(outs)


def _geo_satangles(yr, mo, day, hr, min, sec, latp, lonp, lonss):
    """view_azimuth, view_zenith, solar_azimuth, solar_zenith = satangles(...)   (geostationary: tz 0, latss 0, 35786 km)"""
    n, ints, dbls = _common(yr, mo, day, hr, min, sec, latp, lonp)
    return _call(lib().ab200_geo_satangles, "ab200_geo_satangles", n, ints, dbls, [C.c_double(lonss)], 4)


def _lidar(entry, name):
    def f(yr, mo, day, hr, min, sec, latp, lonp, latss, lonss, tz, hgtss):
        n, ints, dbls = _common(yr, mo, day, hr, min, sec, latp, lonp)
        la, lo = _d(latss, n), _d(lonss, n)
        return _call(getattr(lib(), entry), entry, n, ints, dbls + [la, lo], [C.c_double(tz), C.c_double(hgtss)], 4)
    f.__name__ = name
    return f


def _solarangles(yr, mo, day, hr, min, sec, latp, lonp, tz):
    """solar_azimuth, solar_zenith = solarangles(...)"""
    n, ints, dbls = _common(yr, mo, day, hr, min, sec, latp, lonp)
    return _call(lib().ab200_solarangles, "ab200_solarangles", n, ints, dbls, [C.c_double(tz)], 2)


GeoAngles_ = types.SimpleNamespace(satangles=_geo_satangles)
LidarAngles_ = types.SimpleNamespace(satangles=_lidar("ab200_lidar_satangles", "satangles"),
                                     satangles_fromspace=_lidar("ab200_lidar_satangles_fromspace", "satangles_fromspace"),
                                     solarangles=_solarangles)
stnAngles_ = types.SimpleNamespace(sunangles=_solarangles)
