"""Pins the Rayleigh part of the oracle (VLIDORT_ROT_CALC -> VLIDORT_Rayleigh, V283/VLIDORT_ScatMod.F90:141-281: cross section A(lambda),
King factor F_air, depolarisation ratio, layer optical thickness by trapezoids of the volume coefficient) to vectors produced by the
REFERENCE's own Python restatement of A and F_air (LV/benchmark_polar_vlidort_hyperspectral.py:536-575, executed unmodified by
tests/golden/make_rayleigh_golden.py).  The Fortran writes pi = 4.*atan(1.0), Avogadro's number and the polynomial coefficients as
single-precision literals, which the oracle keeps and the Python does not: the two agree to ~1e-7, and that is the bar here (3e-7)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rayleigh_python_reference.npz")
AV, TS, PS = 6.0221367e23, 288.15, 1013.25


def test_oracle_rayleigh_against_the_reference_python():
    from oracle import vlidort_glue as vg
    g = np.load(GOLD)
    lam_nm = g["lam_um"] * 1e3
    km, nobs = 20, 3
    rng = np.random.default_rng(5)
    pe = np.asfortranarray(np.sort(rng.uniform(1e2, 1.02e5, (km + 1, nobs)), axis=0))            # Pa, increasing downwards
    he = np.asfortranarray(-7500.0 * np.log(pe / 1.02e5) + rng.uniform(0, 50, (1, nobs)))          # m, decreasing downwards
    te = np.asfortranarray(rng.uniform(200.0, 300.0, (km + 1, nobs)))
    ROT, depol = vg.rot_calc(lam_nm, pe, he, te)[:2]
    ROT = np.asarray(ROT); depol = np.asarray(depol)
    assert ROT.shape == (km, nobs, len(lam_nm)) and (ROT > 0).all()
    # King factor back from the depolarisation ratio: depol = 6 (F - 1) / (7 F + 3)
    F = (6.0 + 3.0 * depol) / (6.0 - 7.0 * depol)
    assert np.max(np.abs(F / g["king_factor"] - 1.0)) < 3e-7
    # cross section back from the layer optical thickness: ROT(j+1) = dz_km (Vol(j) + Vol(j+1)) / 2, Vol = A Ns (p/Ps)(Ts/T) 1e5
    Ns = (AV / 22.4141) * (273.15 / TS) * (1.0 / 1000)
    w = (pe * 1e-2 / PS) * (TS / te)
    dz = (he[:-1] - he[1:]) * 1e-3
    geo = dz * 0.5 * (w[:-1] + w[1:]) * Ns * 1e5                                                    # (km, nobs)
    A = ROT / geo[:, :, None]
    assert np.max(np.abs(A / g["cross_section_cm2"][None, None, :] - 1.0)) < 3e-7
    # ... and the differences are the single-precision literals, not noise: one common factor per wavelength
    r = A / g["cross_section_cm2"][None, None, :]
    assert np.max(np.abs(r / r[:1, :1, :] - 1.0)) < 1e-12
