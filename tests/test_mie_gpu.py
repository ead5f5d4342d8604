"""GPU parity tests for Stage A: the CUDA path, called through the C ABI (ctypes mirror MieObs_), against
the CPU oracle on the same seeded inputs.  Bar: BIT-EXACT float32 for tau/ssa/g/pmom and every table
array (north_star: <=1e-12 relative == identical float32); <=1 float32 ulp where the reference calls
libm (exp in getExt's attenuated backscatter, log/pow in getEdgeVars)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from aeroapps_b200 import mietables as mt


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def same(a, b):
    return a.shape == b.shape and np.array_equal(bits(a), bits(b))


def ulp_close(a, b, n=1):
    a = np.asarray(a, np.float32); b = np.asarray(b, np.float32)
    return np.all(np.abs(a - b) <= n * np.spacing(np.maximum(np.abs(a), np.abs(b))))


@pytest.fixture(scope="module")
def env(built):
    from aeroapps_b200 import MieObs_
    from oracle import mie_oracle as orc
    tables = mt.synthetic_tables(nmom=40)
    chan = [354e-9, 550e-9, 600e-9]
    return MieObs_, MieObs_.ChemMie(chan, 40, tables), orc.OracleMie(chan, 40, tables), tables


def test_tables_bit_exact(env):
    _, mie, o, tables = env
    for sp in mt.SPECIES:
        for var in ("rh", "bext", "bsca", "bbck", "g", "refr", "refi", "pback", "pmom", "reff", "gf", "rhop",
                    "rhod", "vol", "area", "rha"):
            assert same(mie.table(sp, var), o.table(sp, var)), (sp, var)
        assert np.array_equal(mie.table(sp, "rhi"), o.table(sp, "rhi")), sp


@pytest.mark.parametrize("nobs,km", [(1, 1), (3, 72), (37, 33), (130, 72)])
def test_aop_scalar_bit_exact(env, nobs, km):
    _, mie, o, _ = env
    st = mt.synthetic_state(nobs=nobs, km=km, seed=nobs + km)
    got = mie.getaopscalar([550.0, 354.0], st["vnames"], 0, st["qm"], st["rh"])
    ref = o.getaopscalar([550.0, 354.0], st["vnames"], st["qm"], st["rh"])
    assert got[-1] == ref[-1] == 0
    for a, b in zip(got[:3], ref[:3]):
        assert same(a, b)


@pytest.mark.parametrize("nobs,km,nmom,npol", [(1, 1, 1, 6), (5, 72, 40, 6), (33, 17, 7, 4), (2, 72, 13, 3)])
def test_aop_vector_bit_exact(env, nobs, km, nmom, npol):
    _, mie, o, _ = env
    st = mt.synthetic_state(nobs=nobs, km=km, seed=7 * nobs + km)
    st["rh"][0, 0] = -0.25                       # clamp paths
    st["qm"][:, 3, :] = 0.0                      # a tracer with no mass
    got = mie.getaopvector([600.0], st["vnames"], 0, st["qm"], st["rh"], nmom, npol)
    ref = o.getaopvector([600.0], st["vnames"], st["qm"], st["rh"], nmom, npol)
    assert got[-1] == ref[-1] == 0
    for a, b in zip(got[:4], ref[:4]):
        assert same(a, b)


def test_spherical_extension_is_order_dependent(env):
    """dust (nPol 6) before/after a spherical tracer changes P22 exactly as in the reference loop"""
    _, mie, o, _ = env
    st = mt.synthetic_state(nobs=4, km=8, vnames=["SO4", "DU001", "SS001"], seed=11)
    for order in ([0, 1, 2], [1, 0, 2], [2, 0, 1]):
        vn = [st["vnames"][i] for i in order]
        qm = np.asfortranarray(st["qm"][:, order, :])
        got = mie.getaopvector([550.0], vn, 0, qm, st["rh"], 10, 6)
        ref = o.getaopvector([550.0], vn, qm, st["rh"], 10, 6)
        assert same(got[3], ref[3])


def test_all_tracers_unknown_and_zero_mass(env):
    _, mie, o, _ = env
    st = mt.synthetic_state(nobs=3, km=5, vnames=["NH3", "NH4a"])
    tau, ssa, g, pmom, rc = mie.getaopvector([550.0], st["vnames"], 0, st["qm"], st["rh"], 4, 6)
    assert rc == 0 and not tau.any() and not ssa.any() and not g.any() and not pmom.any()
    st = mt.synthetic_state(nobs=3, km=5)
    st["qm"][:] = 0
    got = mie.getaopscalar([550.0], st["vnames"], 0, st["qm"], st["rh"])
    ref = o.getaopscalar([550.0], st["vnames"], st["qm"], st["rh"])
    for a, b in zip(got[:3], ref[:3]):
        assert same(a, b)


def test_error_codes_match_reference(env):
    _, mie, o, _ = env
    st = mt.synthetic_state(nobs=2, km=4)
    assert mie.getaopscalar([865.0], st["vnames"], 0, st["qm"], st["rh"])[-1] == 99
    assert mie.getaopvector([550.0], st["vnames"], 0, st["qm"], st["rh"], 41, 6)[-1] == 99
    assert mie.getaopvector([550.0], st["vnames"], 0, st["qm"], st["rh"], 40, 5)[-1] == 777
    assert mie.query_idx("GOCART::du003") == (3, "DU") and mie.query_idx("so2")[0] == -1


def test_ext_int_edge(env):
    MieObs_, mie, o, _ = env
    from oracle import mie_oracle as orc
    st = mt.synthetic_state(nobs=9, km=72, seed=5)
    got = mie.getext([550.0], st["vnames"], 0, st["qc"], st["qm"], st["rh"])
    ref = o.getext([550.0], st["vnames"], st["qc"], st["qm"], st["rh"])
    names = ["ext", "sca", "bsc", "absc_sfc", "absc_toa", "depol"]
    for n, a, b in zip(names, got[:6], ref[:6]):
        if n.startswith("absc"):
            assert ulp_close(a, b, 1), n
        else:
            assert same(a, b), n
    got = mie.getint([550.0], st["vnames"], 0, st["qc"], st["qm"], st["rh"])
    ref = o.getint([550.0], st["vnames"], st["qc"], st["qm"], st["rh"])
    for a, b in zip(got[:5], ref[:5]):
        assert same(a, b)
    pe, ze, te = MieObs_.getedgevars(st["airdens"], st["delp"], 1.0)
    ope, oze, ote = orc.getedgevars(st["airdens"], st["delp"], 1.0)
    assert same(pe, ope) and same(ze, oze) and ulp_close(te, ote, 2)


def test_full_size_c90_properties(env):
    """BASELINE config 1 at full size (48 600 columns x 72 levels): size-independent properties --
    exact linearity in mass under power-of-two scaling, ssa/g invariance, and a spot check of 64
    random columns against the oracle."""
    _, mie, o, _ = env
    st = mt.synthetic_state(nobs=48600, km=72, seed=2026)
    tau, ssa, g, rc = mie.getaopscalar([550.0], st["vnames"], 0, st["qm"], st["rh"])
    assert rc == 0 and np.isfinite(tau).all()
    tau2, ssa2, g2, _ = mie.getaopscalar([550.0], st["vnames"], 0, st["qm"] * np.float32(4), st["rh"])
    assert same(tau2, tau * np.float32(4))
    assert np.max(np.abs(ssa2 - ssa)) < 1e-6 and np.max(np.abs(g2 - g)) < 1e-6
    idx = np.random.default_rng(1).choice(48600, 64, replace=False)
    ref = o.getaopscalar([550.0], st["vnames"], np.asfortranarray(st["qm"][:, :, idx]), np.asfortranarray(st["rh"][:, idx]))
    assert same(np.asfortranarray(tau[:, :, idx]), ref[0]) and same(np.asfortranarray(ssa[:, :, idx]), ref[1])
    assert same(np.asfortranarray(g[:, :, idx]), ref[2])


def test_aop_vector_nmom300_npol6_bit_exact(built):
    """the bench / BASELINE moment count (nMom = 300, nPol = 6: LV/polar_vlidort.py:49, GV/geo_vlidort_pars.F90:14): every one of
    the 1800 moment elements per voxel bit-identical to the oracle, 48 columns x 72 levels, 20 tracers, two channels"""
    from aeroapps_b200 import MieObs_
    from oracle import mie_oracle as orc
    tables = mt.synthetic_tables(nmom=300)
    chan = [388e-9, 550e-9]
    mie, o = MieObs_.ChemMie(chan, 300, tables), orc.OracleMie(chan, 300, tables)
    st = mt.synthetic_state(nobs=48, km=72, seed=300)
    st["rh"][3, 5] = 1.2; st["rh"][0, 0] = -0.1
    got = mie.getaopvector([550.0, 388.0], st["vnames"], 0, st["qm"], st["rh"], 300, 6)
    ref = o.getaopvector([550.0, 388.0], st["vnames"], st["qm"], st["rh"], 300, 6)
    assert got[-1] == ref[-1] == 0
    assert got[3].shape == (72, 2, 48, 300, 6)
    for a, b in zip(got[:4], ref[:4]):
        assert same(a, b)


def test_scalar_handle_with_4_component_tables(built):
    """a handle created with n_moments = 0 from tables whose pback has 4 components (every species but dust)
    must not read past the caller's pback buffer; getext's depolarisation reads P22 through it"""
    from aeroapps_b200 import MieObs_
    from oracle import mie_oracle as orc
    tables = mt.synthetic_tables(nmom=0)
    chan = [550e-9]
    mie, o = MieObs_.ChemMie(chan, 0, tables), orc.OracleMie(chan, 0, tables)
    st = mt.synthetic_state(nobs=7, km=72, seed=44)
    got = mie.getaopscalar([550.0], st["vnames"], 0, st["qm"], st["rh"])
    ref = o.getaopscalar([550.0], st["vnames"], st["qm"], st["rh"])
    for a, b in zip(got[:3], ref[:3]):
        assert same(a, b)
    got = mie.getext([550.0], st["vnames"], 0, st["qc"], st["qm"], st["rh"])
    ref = o.getext([550.0], st["vnames"], st["qc"], st["qm"], st["rh"])
    for n, a, b in zip(["ext", "sca", "bsc", "absc_sfc", "absc_toa", "depol"], got[:6], ref[:6]):
        assert (ulp_close(a, b, 1) if n.startswith("absc") else same(a, b)), n


def test_interpolated_float32_tables_round_trip(env):
    """the Fortran boundary (aeroapps_b200/fortran/ab200_mie_shims.F90) hands over the reference's type(Chem_MieTable) members:
    float32 arrays ALREADY interpolated to the channels by Chem_MieTableRead (CB/Chem_MieTableMod.F90:108-420).  A handle filled
    through ab200_mie_load_table_f32 from those arrays must give bit-identical AOPs to the handle that read the optics files"""
    MieObs_, mie, o, tables = env
    chan = [354e-9, 550e-9, 600e-9]
    m2 = MieObs_.ChemMie(chan, 40, {})
    names = ("rh", "reff", "gf", "rhop", "rhod", "vol", "area", "bext", "bsca", "bbck", "g", "refr", "refi", "pback", "pmom", "rhi", "rha")
    for sp in mt.SPECIES:
        raw = tables[sp]
        m2.load_interpolated(sp, raw.nrh, raw.nbin, raw.npol, {k: mie.table(sp, k) for k in names})
        for k in names:
            assert np.array_equal(m2.table(sp, k).view(np.uint32), mie.table(sp, k).view(np.uint32)), (sp, k)
    st = mt.synthetic_state(nobs=21, km=72, seed=5)
    a = mie.getaopvector([550.0, 354.0], st["vnames"], 0, st["qm"], st["rh"], 40, 6)
    b = m2.getaopvector([550.0, 354.0], st["vnames"], 0, st["qm"], st["rh"], 40, 6)
    for x, y in zip(a[:4], b[:4]):
        assert same(x, y)
    ref = o.getext([550.0], st["vnames"], st["qc"], st["qm"], st["rh"])
    got = m2.getext([550.0], st["vnames"], 0, st["qc"], st["qm"], st["rh"])
    for n, x, y in zip(["ext", "sca", "bsc", "absc_sfc", "absc_toa", "depol"], got[:6], ref[:6]):
        assert (ulp_close(x, y, 1) if n.startswith("absc") else same(x, y)), n
