"""Dimension limits of VLIDORT_Init (V283/VLIDORT_Mod.F90:366-374 with the GEOS-sized maxima of VD/vlidort_pars.f90: MAXSTREAMS 25,
MAXLAYERS 72, MAXMOMENTS_INPUT 1000): the reference returns rc = 2 / 3 / 5 before touching a pixel, and so must the drop-in -- through the
C ABI, before any device work (this test runs without a GPU), outputs untouched."""
import numpy as np


def _args(km, nmom, nobs=2, nch=1):
    z3 = np.zeros((km, nch, nobs), order="F") + 0.1
    pm = np.zeros((km, nch, nobs, nmom, 6), order="F")
    e = np.asfortranarray(np.linspace(1.0, 2.0, (km + 1) * nobs).reshape(km + 1, nobs))
    return dict(ch=np.array([550.0]), ROT=np.zeros((km, nobs, nch), order="F") + 0.01, depol=np.array([0.03]), z3=z3, pm=pm, e=e,
                ang=np.full(nobs, 30.0), w=np.full(nobs, 3.0), mr=np.array([1.334]), alb=np.full((nobs, nch), 0.1, order="F"))


def test_init_limits_are_reported_like_the_reference():
    from aeroapps_b200 import VLIDORT_POLAR_ as VP
    a = _args(73, 4)                                                            # one layer too many
    I, R, rc = VP.scalar_lambert(a["ch"], 6, False, a["ROT"], a["depol"], a["z3"], a["z3"], a["z3"], a["pm"], a["e"], a["e"], a["e"], a["alb"],
                                 a["ang"], a["ang"], a["ang"], 1e15, 0)
    assert rc == 3 and (I == 0).all() and (R == 0).all()
    *out, rc = VP.vector_cx(a["ch"], 6, False, a["ROT"], a["depol"], a["z3"], a["z3"], a["pm"], a["e"], a["e"], a["e"], a["w"], a["w"], a["mr"],
                            a["ang"], a["ang"], a["ang"], 1e15, 0)
    assert rc == 3 and all((o == 0).all() for o in out)
    a = _args(3, 1001)                                                          # one moment too many
    I, R, Q, U, rc = VP.vector_lambert(a["ch"], 6, False, a["ROT"], a["depol"], a["ROT"] * 0, a["z3"], a["z3"], a["pm"], a["e"], a["e"], a["e"],
                                       a["alb"], a["ang"], a["ang"], a["ang"], np.ones((1, 2), order="F"), 1e15, 0)
    assert rc == 5 and (I == 0).all()
    *out, rc = VP.vector_gisscx(a["ch"], 6, False, a["ROT"], a["depol"], a["z3"], a["z3"], a["pm"], a["e"], a["e"], a["e"], a["w"], a["w"], a["mr"],
                                a["ang"], a["ang"], a["ang"], 1e15, 0)
    assert rc == 5
    a = _args(3, 4)                                                             # one stream too many
    *out, rc = VP.vector_brdf_modis(a["ch"], 26, False, a["ROT"], a["depol"], a["z3"], a["z3"], a["pm"], a["e"], a["e"], a["e"],
                                    np.ones((3, 1, 2), order="F"), np.ones((2, 1, 2), order="F"), a["ang"], a["ang"], a["ang"], 1e15, 0)
    assert rc == 2
