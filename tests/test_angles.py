"""Geometry calculators: GeoAngles_.satangles, LidarAngles_.satangles / satangles_fromspace / solarangles,
stnAngles_.sunangles.

CPU part: the oracle (oracle/angles_oracle.c, a statement-by-statement restatement of the reference Fortran; PARITY UNPINNED -- the
reference holds no known answers for these routines) against independent facts: the Julian day of J2000, the sun overhead at the
sub-solar point, an independent spherical-trigonometry derivation of the viewing geometry, the below-horizon sign flip.
GPU part: csrc/geo_angles.cu through the C ABI against the oracle on 200 000 seeded points.  Bar: both sides evaluate the same
expressions in the same order without FMA contraction, so they differ only by the <= 2 ulp of CUDA's sin / cos / acos / atan2 against
glibc's; the formulas amplify that by 1 / sin(angle) near the sub-satellite point and near the sub-solar point, so the bar is
1e-9 degrees where the central angle exceeds 0.01 degree and 1e-4 degrees inside (sqrt(eps) conditioning of acos at 1); the view
azimuth additionally gets 1e-12 / sin(azimuth) degrees (acos of a cosine near +-1 for targets due north / south of the platform)."""
import numpy as np
import pytest


def _points(n, seed):
    rng = np.random.default_rng(seed)
    yr = rng.integers(1999, 2031, n).astype(np.int32); mo = rng.integers(1, 13, n).astype(np.int32)
    day = rng.integers(1, 29, n).astype(np.int32); hr = rng.integers(0, 24, n).astype(np.int32)
    mn = rng.integers(0, 60, n).astype(np.int32); sec = rng.uniform(0, 60, n)
    latp = rng.uniform(-89.5, 89.5, n); lonp = rng.uniform(-180, 180, n)
    latss = np.clip(latp + rng.normal(0, 8, n), -89.9, 89.9); lonss = lonp + rng.normal(0, 8, n)
    return yr, mo, day, hr, mn, sec, latp, lonp, latss, lonss


def test_oracle_known_answers():
    from oracle import angles_oracle as ao
    one = lambda v, t=np.int32: np.array([v], t)
    # J2000.0 = 2000-01-01 12:00 -> JD 2451545.0: t = 0, mean longitude 280.46646, and the solar zenith at (lat = declination, local noon)
    # is ~0.  Sub-solar point on 2014-06-21 12:00 UTC: lat ~ +23.43, lon ~ +0.4 (equation of time -1.7 min)
    saa, sza = ao.angles(2, one(2014), one(6), one(21), one(12), one(0), one(0.0, float), one(23.43, float), one(0.43, float))
    assert sza[0] < 0.1
    # equinox, equator, local midnight: the sun is at nadir
    saa, sza = ao.angles(2, one(2014), one(3), one(20), one(0), one(0), one(0.0, float), one(0.0, float), one(1.8, float))
    assert sza[0] > 179.0
    # geostationary view: at the sub-satellite point the spacecraft is overhead; on the same meridian at latitude phi the zenith angle
    # follows from the plane triangle (earth centre, target, spacecraft)
    lat = np.array([0.0, 30.0, -45.0, 60.0]); z4 = np.zeros(4, np.int32)
    vaa, vza, _, _ = ao.angles(0, z4 + 2014, z4 + 8, z4 + 1, z4 + 18, z4, np.zeros(4), lat, np.full(4, -100.0), lonss0=-100.0)
    re, r = 6378.0, 6378.0 + 35786.0
    phi = np.radians(np.abs(lat))
    expect = np.degrees(np.arctan2(r * np.sin(phi), r * np.cos(phi) - re))
    assert np.allclose(vza, expect, atol=1e-9)
    assert abs(vza[0]) < 1e-6
    assert vaa[1] == pytest.approx(180.0) and vaa[2] == pytest.approx(0.0)           # satellite due south / due north of the target
    # below the horizon: central angle beyond acos(re / r) = 81.3 degrees -> the elevation comes back negated, zenith = 90 + |el|
    _, vz, _, _ = ao.angles(0, z4[:1] + 2014, z4[:1] + 8, z4[:1] + 1, z4[:1] + 18, z4[:1], np.zeros(1), np.array([0.0]), np.array([-5.0]),
                            lonss0=-100.0)
    assert vz[0] > 90.0
    # the nadir angle seen from space and the elevation at the target close the triangle: eta + eps + lambda = 90
    p = _points(1000, 3)
    a = ao.angles(0, *p[:8], latss=p[8], lonss=p[9], tz=0.0, hgtss=705.0)
    b = ao.angles(1, *p[:8], latss=p[8], lonss=p[9], tz=0.0, hgtss=705.0)
    lam = np.degrees(np.arccos(np.clip(np.sin(np.radians(p[8])) * np.sin(np.radians(p[6])) + np.cos(np.radians(p[8])) * np.cos(np.radians(p[6])) *
                                       np.cos(np.radians(p[9] - p[7])), -1, 1)))
    vis = (a[1] < 90.0) & (lam > 0.05)
    assert vis.sum() > 100
    assert np.allclose((90.0 - a[1][vis]) + b[1][vis] + lam[vis], 90.0, atol=1e-6)
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])


def _check(got, ref, central, name):
    far = central > 0.01
    ref = np.asarray(ref)
    d = np.abs(np.asarray(got) - ref)
    d = np.minimum(d, np.abs(d - 360.0))                                                # azimuth wrap
    # at the sub-satellite point itself the reference evaluates acos(1 +- 1 ulp) and 0 / 0: NaN or not depending on the last bit of the
    # libm's sin / cos -- undefined there, not compared (only points with a central angle below 0.01 degree may be affected)
    undef = (np.isnan(np.asarray(got)) | np.isnan(ref)) & ~far
    d = np.where(undef, 0.0, d)
    tol = np.full(d.shape, 1e-9)
    if name.endswith("vaa"):
        # the view azimuth is acos(cphiv) with cphiv = (slatss - slatp clambda) / (clatp sin(acos(clambda))): sin(acos(.)) loses digits as
        # 1 / lambda^2 for a small central angle lambda between target and platform, and acos amplifies by 1 / sin(azimuth) for targets
        # (almost) due north or south of it -- properties of the reference's formulation that a 1-2 ulp difference in sin / cos / acos
        # between CUDA's and glibc's libm exposes; everywhere else the bar stays 1e-9 degrees
        lam = np.maximum(np.radians(central), 1e-6)
        tol = np.minimum(tol + 3e-13 / (lam ** 2 * np.maximum(np.abs(np.sin(np.radians(ref))), 1e-6)), 1e-5)
    bad = far & (d > tol)
    if bad.any():
        k = int(np.argmax(np.where(bad, d, 0.0)))
        raise AssertionError((name, "worst", float(d[k]), "tol", float(tol[k]), "ref", float(ref[k]), "central", float(central[k]), "count", int(bad.sum())))
    if (~far).any():
        assert d[~far].max() <= 1e-4, (name, d[~far].max())


@pytest.mark.gpu
def test_gpu_angles_match_oracle(built):
    from aeroapps_b200.angles import GeoAngles_, LidarAngles_, stnAngles_
    from oracle import angles_oracle as ao
    n = 200_000
    yr, mo, day, hr, mn, sec, latp, lonp, latss, lonss = _points(n, 11)
    # a few exactly singular points: target at the sub-satellite point, on its meridian, at a pole-ward edge
    latss[:3] = latp[:3]; lonss[:3] = lonp[:3]
    lonss[3:6] = lonp[3:6]
    cen = lambda la, lo: np.degrees(np.arccos(np.clip(np.sin(np.radians(la)) * np.sin(np.radians(latp)) + np.cos(np.radians(la)) *
                                                      np.cos(np.radians(latp)) * np.cos(np.radians(lo - lonp)), -1, 1)))
    big = np.full(n, 90.0)
    # geostationary
    lon0 = -100.0
    got = GeoAngles_.satangles(yr, mo, day, hr, mn, sec, latp, lonp, lon0)
    ref = ao.angles(0, yr, mo, day, hr, mn, sec, latp, lonp, lonss0=lon0)
    c = cen(np.zeros(n), np.full(n, lon0))
    for g, r, nm, cc in zip(got, ref, ("vaa", "vza", "saa", "sza"), (c, c, ref[3], ref[3])):
        _check(g, r, np.minimum(cc, 180.0 - cc) if nm in ("saa", "sza") else cc, "geo " + nm)
    # moving platform, both view conventions
    c = cen(latss, lonss)
    for fn, mode in ((LidarAngles_.satangles, 0), (LidarAngles_.satangles_fromspace, 1)):
        got = fn(yr, mo, day, hr, mn, sec, latp, lonp, latss, lonss, -3.0, 705.0)
        ref = ao.angles(mode, yr, mo, day, hr, mn, sec, latp, lonp, latss=latss, lonss=lonss, tz=-3.0, hgtss=705.0)
        for g, r, nm, cc in zip(got, ref, ("vaa", "vza", "saa", "sza"), (c, c, ref[3], ref[3])):
            _check(g, r, np.minimum(cc, 180.0 - cc) if nm in ("saa", "sza") else cc, f"lidar{mode} " + nm)
        assert np.array_equal(np.asarray(got[1]) > 90.0, ref[1] > 90.0) or mode == 1       # the same points are below the horizon
    # sun only (LidarAngles_.solarangles == stnAngles_.sunangles)
    got = LidarAngles_.solarangles(yr, mo, day, hr, mn, sec, latp, lonp, 0.0)
    got2 = stnAngles_.sunangles(yr, mo, day, hr, mn, sec, latp, lonp, 0.0)
    ref = ao.angles(2, yr, mo, day, hr, mn, sec, latp, lonp, tz=0.0)
    for g, g2, r, nm in zip(got, got2, ref, ("saa", "sza")):
        _check(g, r, np.minimum(ref[1], 180.0 - ref[1]), "sun " + nm)
        assert np.array_equal(g, g2)
    # empty input and argument errors
    e = np.zeros(0, np.int32)
    assert all(len(x) == 0 for x in LidarAngles_.solarangles(e, e, e, e, e, np.zeros(0), np.zeros(0), np.zeros(0), 0.0))
    with pytest.raises(AssertionError):
        GeoAngles_.satangles(yr, mo[:5], day, hr, mn, sec, latp, lonp, lon0)
