"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/libmie_oracle.so (the CPU restatement of
Stage A).  Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; never by the product package aeroapps_b200/.  PARITY UNPINNED (see mie_oracle.c header).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmie_oracle.so")
NRHHASH = 991
SPECIES = ["DU", "SS", "BC", "OC", "BRC", "SU", "NI"]
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)


class _Table(C.Structure):
    _fields_ = [("nch", C.c_int), ("nrh", C.c_int), ("nbin", C.c_int), ("nmom", C.c_int), ("npol", C.c_int),
                ("rh", _f32p)] + [(n, _f32p) for n in ("bext", "bsca", "bbck", "g", "refr", "refi", "pback", "pmom",
                                                       "reff", "gf", "rhop", "rhod", "vol", "area")] + \
               [("rhi", C.c_int * NRHHASH), ("rha", C.c_float * NRHHASH)]


class _Raw(C.Structure):
    _fields_ = [("nlam", C.c_int), ("nrh", C.c_int), ("nbin", C.c_int), ("nmom", C.c_int), ("npol", C.c_int)] + \
               [(n, _f64p) for n in ("lam", "rh", "reff", "bext", "bsca", "bbck", "g", "refr", "refi", "pback",
                                     "pmom", "gf", "rhop")]


_lib = None
_native = None
_use_native = False


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def use_native(on=True):
    """bench.py's CPU arms: time a -O3 -march=native build of mie_oracle.c (still -ffp-contract=off: the float32 operation order is the
    algorithm), built on first use on the host that runs it into oracle/_native/.  The portable -O2 build stays the checker of the parity
    tests.  Objects (OracleMie) must be created while the build they are to run with is active.  Returns True when the native build is on."""
    global _native, _use_native
    if on and _native is None:
        from .vlidort_oracle import _cpu_tag
        d = os.path.join(_HERE, "_native")
        so = os.path.join(d, f"libmie_oracle_{_cpu_tag()}.so")
        try:
            if not os.path.exists(so):
                os.makedirs(d, exist_ok=True)
                subprocess.check_call(["gcc", "-O3", "-march=native", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", so,
                                       os.path.join(_HERE, "mie_oracle.c"), "-lm"])
            _native = C.CDLL(so)
            _native.orc_mie_table_read.restype = C.POINTER(_Table)
        except Exception:
            _native = False
    _use_native = bool(on and _native)
    return _use_native


def lib():
    global _lib
    if _use_native:
        return _native
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_mie_table_read.restype = C.POINTER(_Table)
    return _lib


def _p32(a):
    return a.ctypes.data_as(_f32p)


# default registry: Chem_MieRegistry.rc variable tables + the SU/NI filter of CB/Chem_MieMod.F90:343-400
def default_registry():
    reg = {}
    for b in range(1, 6):
        reg["DU%03d" % b] = ("DU", b)
        reg["SS%03d" % b] = ("SS", b)
    reg.update({"BCPHOBIC": ("BC", 1), "BCPHILIC": ("BC", 2), "OCPHOBIC": ("OC", 1), "OCPHILIC": ("OC", 2),
                "BRCPHOBIC": ("BRC", 1), "BRCPHILIC": ("BRC", 2), "SO4": ("SU", 1),
                "NO3AN1": ("NI", 1), "NO3AN2": ("NI", 2), "NO3AN3": ("NI", 3)})
    return reg


class OracleMie:
    """Chem_Mie on the CPU: tables interpolated to `channels_m`, keeping `nmom` moments."""

    def __init__(self, channels_m, nmom, raw_tables):
        self.channels = np.ascontiguousarray(channels_m, dtype=np.float32)
        self.nmom = int(nmom)
        self.tabs = {}
        self._keep = []
        self.reg = default_registry()
        for name, raw in raw_tables.items():
            r = _Raw(raw.nlam, raw.nrh, raw.nbin, raw.nmom, raw.npol)
            for f in ("lam", "rh", "reff", "bext", "bsca", "bbck", "g", "refr", "refi", "pback", "pmom", "gf", "rhop"):
                a = getattr(raw, f)
                setattr(r, f, a.ctypes.data_as(_f64p) if a is not None else None)
            self.tabs[name] = lib().orc_mie_table_read(C.byref(r), len(self.channels), _p32(self.channels), self.nmom)
            self._keep.append(raw)

    def table(self, species, var):
        t = self.tabs[species].contents
        n3 = t.nrh * t.nch * t.nbin
        n2 = t.nrh * t.nbin
        if var == "rhi":
            return np.array(t.rhi[:], dtype=np.int32)
        if var == "rha":
            return np.array(t.rha[:], dtype=np.float32)
        n = {"rh": t.nrh, "pback": n3 * t.npol, "pmom": n3 * t.nmom * t.npol}.get(
            var, n2 if var in ("reff", "gf", "rhop", "rhod", "vol", "area") else n3)
        return np.ctypeslib.as_array(getattr(t, var), shape=(n,)).copy()

    def _resolve(self, vnames, channels):
        nq = len(vnames)
        tabs = (C.POINTER(_Table) * nq)()
        bins = (C.c_int * nq)()
        for i, v in enumerate(vnames):
            v = v.strip().upper()
            if "::" in v:
                v = v.split("::", 1)[1]
            if v in self.reg and self.reg[v][0] in self.tabs:
                tabs[i] = self.tabs[self.reg[v][0]]
                bins[i] = self.reg[v][1]
            else:
                tabs[i] = None
                bins[i] = -1
        ch = np.ascontiguousarray(channels, dtype=np.float32).reshape(-1)
        chidx = (C.c_int * len(ch))()
        for n, c in enumerate(ch):
            idx = -1
            for m, cm in enumerate(self.channels):
                if abs(np.float32(c) - np.float32(1.e9) * cm) < 1.0:      # MieObs_py.F90:148-156
                    idx = m + 1
                    break
            if idx < 0:
                return None, None, None, ch
            chidx[n] = idx
        return tabs, bins, chidx, ch

    def getaopscalar(self, channels, vnames, qm, rh):
        qm = np.asfortranarray(qm, np.float32); rh = np.asfortranarray(rh, np.float32)
        km, nq, nobs = qm.shape
        tabs, bins, chidx, ch = self._resolve(vnames, channels)
        out = [np.zeros((km, len(ch), nobs), np.float32, order="F") for _ in range(3)]
        if tabs is None:
            return (*out, 99)
        rc = lib().orc_get_aop_scalar(km, nobs, len(ch), nq, tabs, bins, chidx, _p32(qm), _p32(rh), *map(_p32, out))
        return (*out, rc)

    def getaopvector(self, channels, vnames, qm, rh, nmom, npol):
        qm = np.asfortranarray(qm, np.float32); rh = np.asfortranarray(rh, np.float32)
        km, nq, nobs = qm.shape
        tabs, bins, chidx, ch = self._resolve(vnames, channels)
        out = [np.zeros((km, len(ch), nobs), np.float32, order="F") for _ in range(3)]
        pmom = np.zeros((km, len(ch), nobs, nmom, npol), np.float32, order="F")
        if tabs is None or nmom > self.nmom:
            return (*out, pmom, 99)
        rc = lib().orc_get_aop_vector(km, nobs, len(ch), nq, tabs, bins, chidx, _p32(qm), _p32(rh), int(nmom),
                                      int(npol), *map(_p32, out), _p32(pmom))
        return (*out, pmom, rc)

    def getext(self, channels, vnames, qc, qm, rh):
        qc = np.asfortranarray(qc, np.float32); qm = np.asfortranarray(qm, np.float32)
        rh = np.asfortranarray(rh, np.float32)
        km, nq, nobs = qm.shape
        tabs, bins, chidx, ch = self._resolve(vnames, channels)
        out = [np.zeros((km, len(ch), nobs), np.float32, order="F") for _ in range(6)]
        if tabs is None:
            return (*out, 99)
        rc = lib().orc_get_ext(km, nobs, len(ch), nq, tabs, bins, chidx, _p32(qc), _p32(qm), _p32(rh), *map(_p32, out))
        return (*out, rc)      # ext, sca, bsc, absc_sfc, absc_toa, depol

    def getint(self, channels, vnames, qc, qm, rh):
        qc = np.asfortranarray(qc, np.float32); qm = np.asfortranarray(qm, np.float32)
        rh = np.asfortranarray(rh, np.float32)
        km, nq, nobs = qm.shape
        tabs, bins, chidx, ch = self._resolve(vnames, channels)
        out = [np.zeros((km, len(ch), nobs), np.float32, order="F") for _ in range(5)]
        if tabs is None:
            return (*out, 99)
        rc = lib().orc_get_int(km, nobs, len(ch), nq, tabs, bins, chidx, _p32(qc), _p32(qm), _p32(rh), *map(_p32, out))
        return (*out, rc)      # vol, area, refr, refi, reff


def getedgevars(airdens, delp, ptop):
    airdens = np.asfortranarray(airdens, np.float32); delp = np.asfortranarray(delp, np.float32)
    km, nobs = airdens.shape
    out = [np.zeros((km + 1, nobs), np.float32, order="F") for _ in range(3)]
    lib().orc_get_edge_vars(km, nobs, _p32(airdens), _p32(delp), C.c_float(ptop), *map(_p32, out))
    return tuple(out)
