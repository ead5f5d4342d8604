"""Host-side mirror of the reference's f2py module ``MieObs_`` (CB/MieObs_py.F90, built by
CB/CMakeLists.txt:22-30): same function names, argument order and returned tuples
(``tau, ssa, g, rc = getaopscalar(rcfile, channels, vname, verbose, qm, rh)`` ...), computed by the
CUDA library through the C ABI of include/aeroapps_b200.h.  No CPU fallback.

``ChemMie`` is the table handle (``Chem_MieCreate``/``Chem_MieDestroy``); the module-level functions
create it from the rc file like the reference does on every call, but keep it cached per rc file.
"""
import ctypes as C
import os
import numpy as np

from ._lib import lib, check
from . import mietables as mt

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)


def _p32(a):
    return a.ctypes.data_as(_f32p)


def _p64(a):
    return None if a is None else a.ctypes.data_as(_f64p)


def _f32(a):
    return np.asfortranarray(a, dtype=np.float32)


def _names(vname):
    """accept f2py-style (nq,16) 'S1' array or a list of str -> contiguous column-major bytes"""
    if isinstance(vname, (list, tuple)):
        vname = mt.pad_names(list(vname))
    a = np.asfortranarray(vname)
    if a.dtype.kind == "U":
        a = a.astype("S1")
    assert a.ndim == 2 and a.shape[1] == 16, "vname must be (nq,16)"
    return a, a.shape[0]


class ChemMie:
    """Chem_Mie handle (CB/Chem_MieMod.F90:65-99) living on the GPU."""

    def __init__(self, channels_m, nmom, tables):
        self._h = C.c_void_p()
        ch = np.ascontiguousarray(channels_m, dtype=np.float32)
        self.channels = ch
        self.nmom = int(nmom)
        check(lib().ab200_mie_create(len(ch), _p32(ch), int(nmom), C.byref(self._h)), "ab200_mie_create")
        for name, raw in tables.items():
            self.load(name, raw)

    @classmethod
    def from_rc(cls, rcfile):
        rc = mt.parse_rc(rcfile)
        tables = {sp: mt.read_optics_file(path) for sp, path in rc["files"].items() if sp in mt.SPECIES}
        return cls(rc["r_channels"], rc["n_moments"], tables)

    def load(self, species, raw):
        sp = mt.SPECIES.index(species)
        rc = lib().ab200_mie_load_table(
            self._h, sp, raw.nlam, raw.nrh, raw.nbin, raw.nmom, raw.npol, _p64(raw.lam), _p64(raw.rh),
            _p64(raw.reff), _p64(raw.bext), _p64(raw.bsca), _p64(raw.bbck), _p64(raw.g), _p64(raw.refr),
            _p64(raw.refi), _p64(raw.pback), _p64(raw.pmom), _p64(raw.gf), _p64(raw.rhop))
        if check(rc, "ab200_mie_load_table") != 0:
            raise ValueError(f"optics table {species}: fewer moments on file than n_moments (rc={rc})")

    def load_interpolated(self, species, nrh, nbin, npol, t):
        """upload an already channel-interpolated float32 table, i.e. the contents of the reference's type(Chem_MieTable)
        (CB/Chem_MieTableMod.F90:39-78) as the Fortran shim passes them; t maps the member names to flat arrays"""
        sp = mt.SPECIES.index(species)
        f = lambda k: (np.ascontiguousarray(t[k], np.float32).ctypes.data_as(C.POINTER(C.c_float)) if t.get(k) is not None else None)
        rhi = np.ascontiguousarray(t["rhi"], np.int32)
        rc = lib().ab200_mie_load_table_f32(
            self._h, sp, int(nrh), int(nbin), int(npol), f("rh"), f("reff"), f("gf"), f("rhop"), f("rhod"), f("vol"), f("area"),
            f("bext"), f("bsca"), f("bbck"), f("g"), f("refr"), f("refi"), f("pback"), f("pmom"),
            rhi.ctypes.data_as(C.POINTER(C.c_int)), f("rha"))
        check(rc, "ab200_mie_load_table_f32")

    def register_tracer(self, name, species, bin_):
        check(lib().ab200_mie_register_tracer(self._h, name.encode(), mt.SPECIES.index(species), int(bin_)), "register")

    def query_idx(self, name):
        sp = C.c_int(-1)
        b = lib().ab200_mie_query_idx(self._h, name.encode(), C.byref(sp))
        return b, (mt.SPECIES[sp.value] if b > 0 else None)

    def table(self, species, var):
        sp = mt.SPECIES.index(species)
        n = lib().ab200_mie_get_table(self._h, sp, var.encode(), None, 0)
        if n < 0:
            raise KeyError(var)
        out = np.empty(n, dtype=np.int32 if var == "rhi" else np.float32)
        lib().ab200_mie_get_table(self._h, sp, var.encode(), out.ctypes.data_as(C.c_void_p), C.c_long(n))
        return out

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            lib().ab200_mie_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the MieObs_ entry points with the table handle passed in (== GV/VLIDORT_MieMod.F90) ----
    def getaopscalar(self, channels, vname, verbose, qm, rh):
        qm, rh = _f32(qm), _f32(rh)
        km, nq, nobs = qm.shape
        ch = np.ascontiguousarray(channels, dtype=np.float32).reshape(-1)
        vn, nqn = _names(vname)
        assert nqn == nq and rh.shape == (km, nobs)
        tau = np.zeros((km, len(ch), nobs), np.float32, order="F")
        ssa = np.zeros_like(tau); g = np.zeros_like(tau)
        rc = C.c_int(0)
        check(lib().ab200_get_aop_scalar(self._h, km, nobs, len(ch), nq, _p32(ch), vn.ctypes.data_as(C.c_char_p),
                                         int(verbose), _p32(qm), _p32(rh), _p32(tau), _p32(ssa), _p32(g),
                                         C.byref(rc)), "ab200_get_aop_scalar")
        return tau, ssa, g, rc.value

    def getaopvector(self, channels, vname, verbose, qm, rh, nmom, npol):
        qm, rh = _f32(qm), _f32(rh)
        km, nq, nobs = qm.shape
        ch = np.ascontiguousarray(channels, dtype=np.float32).reshape(-1)
        vn, nqn = _names(vname)
        assert nqn == nq and rh.shape == (km, nobs)
        tau = np.zeros((km, len(ch), nobs), np.float32, order="F")
        ssa = np.zeros_like(tau); g = np.zeros_like(tau)
        pmom = np.zeros((km, len(ch), nobs, nmom, npol), np.float32, order="F")
        rc = C.c_int(0)
        check(lib().ab200_get_aop_vector(self._h, km, nobs, len(ch), nq, _p32(ch), vn.ctypes.data_as(C.c_char_p),
                                         int(verbose), _p32(qm), _p32(rh), int(nmom), int(npol), _p32(tau),
                                         _p32(ssa), _p32(g), _p32(pmom), C.byref(rc)), "ab200_get_aop_vector")
        return tau, ssa, g, pmom, rc.value

    def getext(self, channels, vname, verbose, qc, qm, rh):
        qc, qm, rh = _f32(qc), _f32(qm), _f32(rh)
        km, nq, nobs = qm.shape
        ch = np.ascontiguousarray(channels, dtype=np.float32).reshape(-1)
        vn, _ = _names(vname)
        outs = [np.zeros((km, len(ch), nobs), np.float32, order="F") for _ in range(6)]
        rc = C.c_int(0)
        check(lib().ab200_get_ext(self._h, km, nobs, len(ch), nq, _p32(ch), vn.ctypes.data_as(C.c_char_p),
                                  int(verbose), _p32(qc), _p32(qm), _p32(rh), *[_p32(o) for o in outs],
                                  C.byref(rc)), "ab200_get_ext")
        ext, sca, bsc, absc_sfc, absc_toa, depol = outs
        return ext, sca, bsc, absc_sfc, absc_toa, depol, rc.value

    def getint(self, channels, vname, verbose, qc, qm, rh):
        qc, qm, rh = _f32(qc), _f32(qm), _f32(rh)
        km, nq, nobs = qm.shape
        ch = np.ascontiguousarray(channels, dtype=np.float32).reshape(-1)
        vn, _ = _names(vname)
        outs = [np.zeros((km, len(ch), nobs), np.float32, order="F") for _ in range(5)]
        rc = C.c_int(0)
        check(lib().ab200_get_int(self._h, km, nobs, len(ch), nq, _p32(ch), vn.ctypes.data_as(C.c_char_p),
                                  int(verbose), _p32(qc), _p32(qm), _p32(rh), *[_p32(o) for o in outs],
                                  C.byref(rc)), "ab200_get_int")
        vol, area, refr, refi, reff = outs
        return vol, area, refr, refi, reff, rc.value


# ------------------------------------------------------------------------------ f2py-named functions
_cache = {}


def _tables_for(rcfile):
    key = (os.path.abspath(rcfile), os.path.getmtime(rcfile))
    if key not in _cache:
        _cache.clear()
        _cache[key] = ChemMie.from_rc(rcfile)
    return _cache[key]


last_error = ""


def getmiedims(mieTableFile):
    """nCh, nRh, nBin, nMom, nPol, rc  (MieObs_py.F90:7-23 -> Chem_MieTableGetDims)"""
    try:
        raw = mt.read_optics_file(mieTableFile)
    except Exception as e:                                   # the reference returns rc /= 0; keep the reason where a caller can find it
        global last_error
        last_error = f"getmiedims({mieTableFile}): {type(e).__name__}: {e}"
        return 0, 0, 0, 0, 0, 1
    return raw.nlam, raw.nrh, raw.nbin, raw.nmom, raw.npol, 0


def getedgevars(airdens, delp, ptop):
    """pe, ze, te = getedgevars(airdens, delp, ptop)  (MieObs_py.F90:27-83)"""
    airdens, delp = _f32(airdens), _f32(delp)
    km, nobs = airdens.shape
    pe = np.zeros((km + 1, nobs), np.float32, order="F")
    ze = np.zeros_like(pe); te = np.zeros_like(pe)
    check(lib().ab200_get_edge_vars(km, nobs, _p32(airdens), _p32(delp), C.c_float(ptop), _p32(pe), _p32(ze),
                                    _p32(te)), "ab200_get_edge_vars")
    return pe, ze, te


def getaopscalar(rcfile, channels, vname, verbose, qm, rh):
    return _tables_for(rcfile).getaopscalar(channels, vname, verbose, qm, rh)


def getaopvector(rcfile, channels, vname, verbose, qm, rh, nmom, npol):
    return _tables_for(rcfile).getaopvector(channels, vname, verbose, qm, rh, nmom, npol)


def getext(rcfile, channels, vname, verbose, qc, qm, rh):
    return _tables_for(rcfile).getext(channels, vname, verbose, qc, qm, rh)


def getint(rcfile, channels, vname, verbose, qc, qm, rh):
    return _tables_for(rcfile).getint(channels, vname, verbose, qc, qm, rh)
