"""Optics-table handling on the host side of Stage A.

* ``RawTable``      -- the arrays ``Chem_MieTableRead`` pulls out of an optics file
                       (CB/Chem_MieTableMod.F90:238-432), kept as float64 in the file's Fortran order.
* ``read_optics_file`` -- NetCDF-3 (classic) reader via scipy, or ``.npz`` written by ``RawTable.save``.
* ``synthetic_tables`` -- the analytic GOCART-like tables of SURVEY.md section 8(d) (no network, no ExtData).
* ``parse_rc``      -- the keys ``Chem_MieCreateFromRC`` reads (CB/Chem_MieMod.F90:207-296).
"""
import os
import re
import numpy as np

SPECIES = ["DU", "SS", "BC", "OC", "BRC", "SU", "NI"]          # order of include/aeroapps_b200.h

# tracer names of CB/mieobs.py:10-17
VNAMES = ['DU001', 'DU002', 'DU003', 'DU004', 'DU005',
          'SS001', 'SS002', 'SS003', 'SS004', 'SS005',
          'BCPHOBIC', 'BCPHILIC', 'OCPHOBIC', 'OCPHILIC', 'SO4',
          'NH3', 'NH4a', 'NO3an1', 'NO3an2', 'NO3an3', 'BRCphobic', 'BRCphilic']


class RawTable:
    """One optics file.  Arrays are float64, Fortran order as nf_get_var_double returns them:
    lambda(nlam) rh(nrh) reff(nrh,nbin) bext|bsca|bbck|g|refr|refi(nlam,nrh,nbin)
    pback(nlam,nrh,nbin,npol) pmom(nlam,nrh,nbin,nmom,npol) gf|rhop(nrh,nbin)."""

    fields = ["lam", "rh", "reff", "bext", "bsca", "bbck", "g", "refr", "refi", "pback", "pmom", "gf", "rhop"]

    def __init__(self, **kw):
        for f in self.fields:
            v = kw.get(f)
            setattr(self, f, None if v is None else np.asfortranarray(v, dtype=np.float64))
        self.nlam, self.nrh, self.nbin = self.bext.shape
        if self.pmom is not None:
            self.nmom, self.npol = self.pmom.shape[3:]
        else:
            self.nmom, self.npol = 0, (self.pback.shape[3] if self.pback is not None else 6)

    def save(self, path):
        np.savez(path, **{f: getattr(self, f) for f in self.fields if getattr(self, f) is not None})


def read_optics_file(path):
    """Read an optics table: ``.npz`` (RawTable.save) or NetCDF classic (scipy.io.netcdf_file).
    NetCDF dims are C-ordered (nPol,nMom,radius,rh,lambda) == the Fortran order reversed."""
    if path.endswith(".npz"):
        z = np.load(path)
        return RawTable(**{k: z[k] for k in z.files})
    with open(path, "rb") as f:
        magic = f.read(8)
    if magic[:4] == b"\x89HDF":
        # NetCDF-4 / HDF5 container (nf_open reads both): needs netCDF4 or h5py, neither of which ships in this image
        return _read_optics_hdf5(path)
    if magic[:3] != b"CDF":
        raise ValueError(f"{path}: not a NetCDF classic (CDF-1/2/5) or NetCDF-4/HDF5 optics file (magic {magic[:4]!r})")
    from scipy.io import netcdf_file
    with netcdf_file(path, "r", mmap=False) as nc:
        def var(name, default=None):
            if name not in nc.variables:
                return default
            return np.array(nc.variables[name][:], dtype=np.float64).T   # C (reversed dims) -> Fortran order
        kw = dict(lam=var("lambda"), rh=var("rh"), reff=var("rEff"), bext=var("bext"), bsca=var("bsca"),
                  bbck=var("bbck"), g=var("g"), refr=var("refreal"), refi=var("refimag"), pback=var("pback"),
                  pmom=var("pmom"), gf=var("growth_factor"), rhop=var("rhop"))
    return RawTable(**kw)


def _read_optics_hdf5(path):
    """NetCDF-4 optics files through netCDF4 or h5py when one of them is importable; a clear error otherwise"""
    try:
        import netCDF4
        ds = netCDF4.Dataset(path)
        get = lambda n: (np.array(ds.variables[n][:], dtype=np.float64).T if n in ds.variables else None)
    except ImportError:
        try:
            import h5py
        except ImportError:
            raise RuntimeError(f"{path} is a NetCDF-4/HDF5 optics file; reading it needs the netCDF4 or h5py package "
                               "(convert with `nccopy -k classic` or install one of them)") from None
        ds = h5py.File(path, "r")
        get = lambda n: (np.array(ds[n][...], dtype=np.float64).T if n in ds else None)
    try:
        kw = dict(lam=get("lambda"), rh=get("rh"), reff=get("rEff"), bext=get("bext"), bsca=get("bsca"), bbck=get("bbck"), g=get("g"),
                  refr=get("refreal"), refi=get("refimag"), pback=get("pback"), pmom=get("pmom"), gf=get("growth_factor"), rhop=get("rhop"))
    finally:
        ds.close()
    return RawTable(**kw)


def parse_rc(path):
    """n_channels / n_moments / r_channels / filename_optical_properties_XX of an Aod_EOS.rc-style file."""
    out = {"files": {}}
    with open(path) as f:
        for line in f:
            line = line.split("#", 1)[0].strip()
            m = re.match(r"^([A-Za-z_0-9]+):\s*(.*)$", line)
            if not m:
                continue
            key, val = m.group(1), m.group(2).strip()
            if key == "n_channels":
                out["n_channels"] = int(val)
            elif key == "n_moments":
                out["n_moments"] = int(val)
            elif key == "r_channels":
                out["r_channels"] = [float(x) for x in val.replace(",", " ").split()]
            elif key.startswith("filename_optical_properties_"):
                out["files"][key[len("filename_optical_properties_"):]] = val
    out.setdefault("n_moments", 0)
    if "n_channels" in out and "r_channels" in out:
        out["r_channels"] = out["r_channels"][:out["n_channels"]]
    base = os.path.dirname(os.path.abspath(path))
    out["files"] = {k: (v if os.path.isabs(v) else os.path.join(base, v)) for k, v in out["files"].items()}
    return out


# ---------------------------------------------------------------------------------------------
# synthetic GOCART-like tables
_RH = np.concatenate([np.arange(0, 0.8001, 0.05), np.arange(0.81, 0.9901, 0.01)])     # 17 + 19 = 36
_LAM_NM = np.array([300, 354, 388, 470, 550, 670, 870, 1020, 1640, 2130], dtype=np.float64)
# species: nbin, npol, b0 [m2/kg], angstrom, ssa, g, hygroscopic
_SPEC = {
    "DU": (5, 6, [2.9, 1.2, 0.55, 0.31, 0.16], [0.3, 0.1, 0.0, -0.05, -0.1], 0.92, 0.72, False),
    "SS": (5, 4, [3.5, 2.2, 0.9, 0.35, 0.12], [1.2, 0.6, 0.2, 0.0, -0.05], 0.999, 0.78, True),
    "BC": (2, 4, [9.3, 9.3], [1.1, 1.1], 0.25, 0.34, True),
    "OC": (2, 4, [4.0, 4.0], [1.7, 1.7], 0.96, 0.62, True),
    "BRC": (2, 4, [4.2, 4.2], [2.2, 2.2], 0.88, 0.62, True),
    "SU": (1, 4, [3.2], [1.9], 0.999, 0.68, True),
    "NI": (3, 4, [3.0, 1.1, 0.4], [1.8, 0.7, 0.1], 0.999, 0.70, True),
}


def synthetic_table(species, nmom=300):
    nbin, npol, b0, ang, w0, g0, hyg = _SPEC[species]
    nlam, nrh = len(_LAM_NM), len(_RH)
    lam = _LAM_NM * 1e-9
    rh = _RH.copy()
    growth = (1.0 + 3.0 * rh ** 4) if hyg else np.ones_like(rh)
    bext = np.empty((nlam, nrh, nbin), order="F")
    bsca = np.empty_like(bext); bbck = np.empty_like(bext); g = np.empty_like(bext)
    refr = np.empty_like(bext); refi = np.empty_like(bext)
    reff = np.empty((nrh, nbin), order="F"); gf = np.empty_like(reff); rhop = np.empty_like(reff)
    pback = np.empty((nlam, nrh, nbin, npol), order="F")
    pmom = np.empty((nlam, nrh, nbin, nmom, npol), order="F") if nmom > 0 else None
    ll = np.arange(nmom, dtype=np.float64)
    for j in range(nbin):
        r_dry = 0.1e-6 * (2.5 ** j) * (3.0 if species in ("DU", "SS") else 1.0)
        for i in range(nrh):
            gfac = growth[i] ** (1.0 / 3.0)
            reff[i, j] = r_dry * gfac
            gf[i, j] = gfac
            rhop[i, j] = (2650.0 if species == "DU" else 1800.0) / gfac ** 3 + 1000.0 * (1 - 1 / gfac ** 3)
            for l in range(nlam):
                spec = (550.0 / _LAM_NM[l]) ** ang[j]
                bext[l, i, j] = b0[j] * growth[i] * spec
                w = min(0.9999, w0 + (1 - w0) * 0.6 * rh[i] ** 2) * (1.0 - 0.03 * (550.0 / _LAM_NM[l] - 1.0) * (1 - w0))
                w = min(max(w, 0.05), 0.9999)
                bsca[l, i, j] = bext[l, i, j] * w
                gg = min(0.9, max(0.05, g0 * (1 + 0.08 * rh[i]) * (550.0 / _LAM_NM[l]) ** 0.08 * (0.9 + 0.04 * j)))
                g[l, i, j] = gg
                bbck[l, i, j] = bsca[l, i, j] * (1 - gg) ** 2 / (4 * np.pi * (1 + gg) ** 2) * 2.0
                refr[l, i, j] = 1.33 + 0.2 / gfac ** 3
                refi[l, i, j] = -(1 - w) * 0.05          # negative on file; the reader takes abs()
                pb = [1.0 - 0.3 * gg, -0.02 * gg, -(1.0 - 0.3 * gg), 0.01, 0.9 - 0.3 * gg, -(0.9 - 0.3 * gg)]
                pback[l, i, j, :] = pb[:npol]
                if nmom > 0:
                    p11 = (2 * ll + 1) * gg ** ll                      # Henyey-Greenstein
                    ratios = [1.0, -0.12, 0.95, 0.08, 0.97, 0.93]
                    for p in range(npol):
                        col = p11 * ratios[p]
                        if p in (1, 3):
                            col = col * (ll >= 2) * (1.0 - np.exp(-ll / 6.0))
                        pmom[l, i, j, :, p] = col
    return RawTable(lam=lam, rh=rh, reff=reff, bext=bext, bsca=bsca, bbck=bbck, g=g, refr=refr, refi=refi,
                    pback=pback, pmom=pmom, gf=gf, rhop=rhop)


def synthetic_tables(nmom=300, species=SPECIES):
    return {s: synthetic_table(s, nmom) for s in species}


def pad_names(vnames):
    """``character vname(nq,16)`` column-major bytes, as f2py passes ``pad(vtypes)`` (CB/mieobs.py)."""
    nq = len(vnames)
    a = np.full((nq, 16), b" ", dtype="S1", order="F")
    for i, v in enumerate(vnames):
        for c, ch in enumerate(v[:16]):
            a[i, c] = ch.encode()
    return a


def synthetic_state(nobs, km=72, vnames=None, seed=20260102, ptop=1.0):
    """Synthetic GEOS aerosol columns: returns dict with float32 Fortran arrays
    qm(km,nq,nobs) [kg/m2], qc(km,nq,nobs) [kg/m3], rh(km,nobs), delp, airdens (km,nobs)."""
    vnames = vnames or [v for v in VNAMES if v not in ("NH3", "NH4a")]
    nq = len(vnames)
    rng = np.random.default_rng(seed)
    ps = rng.uniform(850e2, 1030e2, nobs)
    # 72-level hybrid-like grid: smooth, monotone in pressure
    eta = np.linspace(0, 1, km + 1) ** 2.2
    pe = ptop + (ps[None, :] - ptop) * eta[:, None]
    delp = np.diff(pe, axis=0)
    pm = 0.5 * (pe[1:] + pe[:-1])
    T = 210.0 + 80.0 * (pm / ps[None, :]) ** 0.9
    airdens = pm / (287.0 * T)
    rh = rng.uniform(0.0, 1.05, (km, nobs))
    mu = {"DU": -19.5, "SS": -20.0, "BC": -22.5, "OC": -21.0, "SO": -20.5, "NO": -21.5, "BR": -22.0, "NH": -22.0}
    q = np.empty((km, nq, nobs))
    prof = np.exp(-((1 - pm / ps[None, :]) * 6.0))      # aerosol concentrated near the surface
    for n, v in enumerate(vnames):
        q[:, n, :] = np.exp(rng.normal(mu.get(v[:2].upper(), -21.0), 1.0, (km, nobs))) * prof
    qm = q * delp[:, None, :] / 9.81
    qc = q * airdens[:, None, :]
    f = lambda a: np.asfortranarray(a, dtype=np.float32)
    return dict(vnames=vnames, qm=f(qm), qc=f(qc), rh=f(rh), delp=f(delp), airdens=f(airdens), ptop=np.float32(ptop))
