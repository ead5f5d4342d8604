"""TEST INFRASTRUCTURE ONLY -- helpers that run the wrapper-level oracle on many pixels with a thread pool and
compare a product result with it.  Used by tests/ (the -m gpu parity tests at the BASELINE shapes) and by the
cpu_baseline / parity leg of bench.py; never imported by the product package.

Bars (BASELINE.json north_star): Stokes I and reflectance <= 1e-8 relative, Q and U <= 1e-8 relative to I, identical
Fourier counts (FOURIER_SAVED, VM/vlidort_converge.f90:1139-1156).  The reference stops its eigenvalue iteration at 1e-12
(ASYMTX, VM/vlidort_solutions.f90:530-575), which leaves up to ~1e-8 of iteration noise in its own radiances for 36 x 36
and 48 x 48 layer problems; a pixel that differs from the reference-tolerance oracle by more than 1e-8 must therefore agree
with the converged oracle (tolerance 1e-16) to 1e-9, and such pixels are counted and reported."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import vlidort_glue as vg

KEYS = ("radiance", "reflectance", "Q", "U", "BR", "BR_Q", "BR_U", "nfourier", "status")


def _run(kind, S, N, planpar, sc, obs_range, scalar_glue, MISSING):
    if kind == "lambert":
        return vg.vlidort_run(S, N, planpar, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["he"], sc["sza"], sc["raa"],
                              sc["vza"], albedo=sc["albedo"], alpha=sc.get("alpha"), flux_factor=sc.get("flux"), MISSING=MISSING,
                              scalar_glue=scalar_glue, obs_range=obs_range)
    return vg.vlidort_run_surface(kind, S, N, planpar, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["he"], sc["sza"],
                                  sc["raa"], sc["vza"], sc.get("u10m"), sc.get("v10m"), sc.get("mr"), sc.get("kernel_wt"), sc.get("param"),
                                  MISSING=MISSING, scalar_glue=scalar_glue, obs_range=obs_range, alpha=sc.get("alpha"),
                                  flux_factor=sc.get("flux"), bpdf=sc.get("bpdf"), albedo=sc.get("albedo"))


def oracle_parallel(kind, S, N, planpar, sc, nthreads=None, scalar_glue=False, MISSING=1e15, obs=None, asymtx_tol=1e-12):
    """wrapper-level oracle over all observations of the scene (or the list `obs`), spread over host threads.
    Returns dict of (nobs, nch) arrays (rows outside `obs` stay zero)."""
    nobs = sc["tau"].shape[2]
    nthreads = nthreads or (os.cpu_count() or 1)
    if obs is None:
        edges = np.linspace(0, nobs, min(nthreads * 4, nobs) + 1).astype(int)
        ranges = [(int(a), int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a]
    else:
        ranges = [(int(j), int(j) + 1) for j in obs]
    vg.set_asymtx_tolerance(asymtx_tol)
    if "tauI" in sc:
        vg.set_cloud(sc)                                          # cloud variants: mixed into the glue of every run below
    try:
        with ThreadPoolExecutor(nthreads) as ex:
            parts = list(ex.map(lambda ab: _run(kind, S, N, planpar, sc, ab, scalar_glue, MISSING), ranges))
    finally:
        vg.set_asymtx_tolerance(1e-12)
        vg.set_cloud(None)
    out = {}
    for k in KEYS:
        if k not in parts[0]:
            continue
        full = np.zeros_like(parts[0][k])
        for (a, b), pr in zip(ranges, parts):
            full[a:b] = pr[k][a:b]
        out[k] = full
    return out


def compare(got, ref, S, MISSING=1e15):
    """got / ref: dicts with radiance, reflectance, Q, U (nobs, nch) [+ nfourier].  Returns per-pixel worst relative error
    (I, reflectance; Q, U relative to I) over the pixels both sides solved, and the bookkeeping masks."""
    Ir = ref["radiance"]
    special = (Ir == MISSING) | (Ir == -500.0)
    same_special = bool(np.array_equal(got["radiance"] == MISSING, Ir == MISSING) and
                        np.array_equal(got["radiance"] == -500.0, Ir == -500.0))
    ok = ~special
    den = np.where(ok, np.abs(Ir), 1.0)
    e_I = np.where(ok, np.abs(got["radiance"] - Ir) / den, 0.0)
    rr = ref["reflectance"]
    e_R = np.where(ok, np.abs(got["reflectance"] - rr) / np.where(ok, np.abs(rr), 1.0), 0.0)
    e_QU = np.zeros_like(e_I)
    if S > 1:
        e_QU = np.where(ok, np.maximum(np.abs(got["Q"] - ref["Q"]), np.abs(got["U"] - ref["U"])) / den, 0.0)
    err = np.maximum(np.maximum(e_I, e_R), e_QU)
    nf_bad = np.zeros_like(ok)
    if "nfourier" in got and "nfourier" in ref:
        nf_bad = ok & (np.asarray(got["nfourier"]).reshape(Ir.shape) != ref["nfourier"])
    return dict(err=err, e_I=e_I, e_QU=e_QU, ok=ok, same_special=same_special, nf_bad=nf_bad)


def parity_report(kind, S, N, planpar, sc, got, nthreads=None, scalar_glue=False, MISSING=1e15, tol=1e-8, tol_conv=1e-9):
    """Full protocol: reference-tolerance oracle on every pixel; pixels above `tol` are re-run with the converged oracle and
    must agree to `tol_conv`.  Returns a JSON-able dict (the `parity` object of the bench line)."""
    ref = oracle_parallel(kind, S, N, planpar, sc, nthreads, scalar_glue, MISSING)
    c = compare(got, ref, S, MISSING)
    over = np.argwhere(c["err"] > tol)
    rep = dict(n=int(c["ok"].sum()), max_rel_I=float(c["e_I"].max()), max_rel_QU_over_I=float(c["e_QU"].max()),
               max_rel=float(c["err"].max()), n_over_1e_8=int(len(over)), n_fourier_mismatch=int(c["nf_bad"].sum()),
               special_pixels_identical=c["same_special"], n_special=int((~c["ok"]).sum()),
               oracle="oracle/ restatement of the reference at its own ASYMTX tolerance (1e-12)")
    if len(over):
        obs = sorted({int(j) for j, _ in over})
        ref2 = oracle_parallel(kind, S, N, planpar, sc, nthreads, scalar_glue, MISSING, obs=obs, asymtx_tol=1e-16)
        sel = np.zeros_like(c["ok"]); sel[obs] = True
        g2 = {k: np.where(sel, got[k], ref2[k]) for k in ("radiance", "reflectance", "Q", "U") if k in got and got[k] is not None}
        c2 = compare(g2, ref2, S, MISSING)
        worst = float(np.max(c2["err"][sel & c["ok"]])) if (sel & c["ok"]).any() else 0.0
        rep.update(max_rel_vs_converged_oracle_on_those=worst, n_over_1e_9_vs_converged=int(((c2["err"] > tol_conv) & sel).sum()))
    else:
        rep.update(max_rel_vs_converged_oracle_on_those=0.0, n_over_1e_9_vs_converged=0)
    rep["pass"] = bool(rep["n_over_1e_9_vs_converged"] == 0 and rep["n_fourier_mismatch"] == 0 and rep["special_pixels_identical"])
    return rep
