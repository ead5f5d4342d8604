"""CPU tests of the Stage-A oracle (no GPU): the C restatement (oracle/mie_oracle.c) is cross-checked
bit-for-bit against an independent numpy float32 restatement of the same reference lines
(CB/Chem_MieTableMod.F90:612-627,671-717; CB/Chem_MieMod.F90:860-945; CB/MieObs_py.F90:173-205,391-446)
on small cases, plus analytic known answers.  The reference holds no golden vectors for Stage A
 -- parity unpinned."""
import numpy as np
import pytest

from aeroapps_b200 import mietables as mt
from oracle import mie_oracle as orc

f32 = np.float32


def np_polint(x, y, xw):
    xw = f32(xw)
    n = len(x)
    i = 0
    if np.float64(xw) < x[0]:
        i = 1
    if np.float64(xw) > x[-1]:
        i = n
    if i == 0:
        for j in range(1, n + 1):
            if np.float64(xw) >= x[j - 1]:
                i = j
    slope = f32(0) if i == n else f32((y[i] - y[i - 1]) / (x[i] - x[i - 1]))
    dx = f32(np.float64(xw) - x[i - 1])
    return f32(y[i - 1] + np.float64(f32(slope * dx)))


def np_hash(rh):
    rhi = np.ones(991, np.int32); rha = np.zeros(991, f32)
    nrh = len(rh)
    for j in range(1, 992):
        for i in range(nrh, 0, -1):
            if (j - 1) >= int(f32(rh[i - 1] * f32(1000))):
                rhi[j - 1] = i
                if i + 1 > nrh:
                    rha[j - 1] = 0
                else:
                    rha[j - 1] = f32(f32(f32(j - 1) / f32(1000)) - rh[i - 1]) / f32(rh[i] - rh[i - 1])
                break
    return rhi, rha


@pytest.fixture(scope="module")
def setup():
    tables = mt.synthetic_tables(nmom=12)
    chan = [470e-9, 550e-9, 600e-9]          # 600 nm is off-grid -> genuine interpolation
    o = orc.OracleMie(chan, 12, tables)
    return tables, chan, o


def test_polint_and_hash_match_numpy(setup):
    tables, chan, o = setup
    raw = tables["DU"]
    bext = o.table("DU", "bext").reshape((raw.nrh, len(chan), raw.nbin), order="F")
    for i in (0, 7, 35):
        for n in range(len(chan)):
            for j in (0, 4):
                assert bext[i, n, j] == np_polint(raw.lam, raw.bext[:, i, j], f32(chan[n]))
    pm = o.table("DU", "pmom").reshape((raw.nrh, len(chan), raw.nbin, 12, raw.npol), order="F")
    assert pm[3, 2, 1, 5, 4] == np_polint(raw.lam, raw.pmom[:, 3, 1, 5, 4], f32(chan[2]))
    rhi, rha = np_hash(o.table("DU", "rh"))
    assert np.array_equal(rhi, o.table("DU", "rhi"))
    assert np.array_equal(rha.view(np.uint32), o.table("DU", "rha").view(np.uint32))
    # refi sign: reader stores abs()
    assert (o.table("DU", "refi") >= 0).all()
    # out-of-range channel clamps (slope 0 at the top, extrapolates with slope at the bottom edge index 1)
    o2 = orc.OracleMie([3000e-9], 0, {"SU": tables["SU"]})
    assert o2.table("SU", "bext")[0] == f32(tables["SU"].bext[-1, 0, 0])


def np_query(o, sp, b, ch, rh):
    """Chem_MieQueryByInt for arrays of rh (float32), returns bext,bsca,g and (irh, irhp1, arh)"""
    T = o.tabs[sp].contents
    rhi = o.table(sp, "rhi"); rha = o.table(sp, "rha")
    rhu = np.minimum(rh, f32(0.99))
    isnap = ((rhu + f32(0.001)) * f32(1000)).astype(np.int32)
    isnap = np.maximum(isnap, 1)
    arh = rha[isnap - 1]; irh = rhi[isnap - 1]; irhp1 = np.minimum(irh + 1, T.nrh)
    def tab(name):
        a = o.table(sp, name).reshape((T.nrh, T.nch, T.nbin), order="F")[:, ch, b - 1]
        return a[irh - 1] * (f32(1) - arh) + a[irhp1 - 1] * arh
    return tab("bext"), tab("bsca"), tab("g"), irh, irhp1, arh


def test_aop_scalar_and_vector_match_numpy(setup):
    tables, chan, o = setup
    st = mt.synthetic_state(nobs=5, km=9, seed=3)
    vn = st["vnames"] + ["NH3"]                       # unknown tracer is skipped (idxTable == -1)
    qm = np.concatenate([st["qm"], np.ones((9, 1, 5), f32)], axis=1)
    rh = st["rh"]
    tau, ssa, g, pmom, rc = o.getaopvector([600.0, 470.0], vn, qm, rh, 12, 6)
    assert rc == 0
    ts, ss, gs, rcs = o.getaopscalar([600.0, 470.0], vn, qm, rh)
    assert rcs == 0
    assert np.array_equal(tau, ts) and np.array_equal(ssa, ss) and np.array_equal(g, gs)
    reg = orc.default_registry()
    for n, chn in enumerate((2, 0)):
        t = np.zeros((9, 5), f32); s = np.zeros((9, 5), f32); gg = np.zeros((9, 5), f32)
        pm = np.zeros((9, 5, 12, 6), f32)
        for iq, v in enumerate(vn):
            if v.upper() not in reg:
                continue
            sp, b = reg[v.upper()]
            T = o.tabs[sp].contents
            bext, bsca, gas, irh, irhp1, arh = np_query(o, sp, b, chn, rh)
            tau_ = bext * qm[:, iq, :]; ssa_ = bsca / bext
            t = t + tau_; s = s + ssa_ * tau_; gg = gg + (gas * ssa_) * tau_
            tp = o.table(sp, "pmom").reshape((T.nrh, T.nch, T.nbin, 12, T.npol), order="F")[:, chn, b - 1]
            for p in range(min(6, T.npol)):
                for m in range(12):
                    a = tp[:, m, p]
                    val = a[irh - 1] * (f32(1) - arh) + a[irhp1 - 1] * arh
                    pm[:, :, m, p] = pm[:, :, m, p] + (val * ssa_) * tau_
            if T.npol == 4:
                pm[:, :, :, 4] = pm[:, :, :, 0]; pm[:, :, :, 5] = pm[:, :, :, 2]
        s = np.where(t > 0, s / np.where(t > 0, t, 1), s)
        d = s * t
        gg = np.where(d > 0, gg / np.where(d > 0, d, 1), gg)
        pm = np.where((d > 0)[:, :, None, None], pm / np.where(d > 0, d, 1)[:, :, None, None], pm)
        assert np.array_equal(t.view(np.uint32), tau[:, n, :].view(np.uint32))
        assert np.array_equal(s.view(np.uint32), ssa[:, n, :].view(np.uint32))
        assert np.array_equal(gg.view(np.uint32), g[:, n, :].view(np.uint32))
        assert np.array_equal(pm.view(np.uint32), pmom[:, n, :, :, :].view(np.uint32))
    # physical sanity: 0 < ssa < 1, P11 moment 0 == 1 for HG tables
    assert (ssa > 0).all() and (ssa < 1).all()
    np.testing.assert_allclose(pmom[:, :, :, 0, 0], 1.0, rtol=2e-6)


def test_error_codes(setup):
    tables, chan, o = setup
    st = mt.synthetic_state(nobs=2, km=4)
    assert o.getaopscalar([865.0], st["vnames"], st["qm"], st["rh"])[-1] == 99       # channel not in tables
    assert o.getaopvector([550.0], st["vnames"], st["qm"], st["rh"], 13, 6)[-1] == 99  # nMom > n_moments
    assert o.getaopvector([550.0], st["vnames"], st["qm"], st["rh"], 12, 5)[-1] == 777  # nPol 5 vs table 4


def test_edge_vars_hydrostatic():
    st = mt.synthetic_state(nobs=3, km=72)
    pe, ze, te = orc.getedgevars(st["airdens"], st["delp"], 1.0)
    assert pe.shape == (73, 3) and np.all(np.diff(pe, axis=0) > 0) and np.all(np.diff(ze, axis=0) < 0)
    assert abs(ze[-1]).max() == 0 and 150 < te.min() and te.max() < 330


def test_ext_int_sanity(setup):
    tables, chan, o = setup
    st = mt.synthetic_state(nobs=3, km=6)
    ext, sca, bsc, sfc, toa, dep, rc = o.getext([550.0], st["vnames"], st["qc"], st["qm"], st["rh"])
    assert rc == 0 and (ext >= sca).all() and (toa <= bsc * 1.0000001).all() and (sfc <= bsc * 1.0000001).all()
    vol, area, refr, refi, reff, rc = o.getint([550.0], st["vnames"], st["qc"], st["qm"], st["rh"])
    assert rc == 0 and (refr > 1.3).all() and (refi >= 0).all() and (reff > 0).all()
