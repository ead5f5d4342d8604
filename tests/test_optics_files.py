"""Host-side optics-file reader (aeroapps_b200/mietables.py::read_optics_file, the stand-in for the nf_open / nf_get_var_double calls of
CB/Chem_MieTableMod.F90:238-420): NetCDF classic round trip in the file's own dimension order, clear errors for NetCDF-4 files when no
HDF5 reader is installed and for files that are not NetCDF at all, and getmiedims keeping the reason of a failure."""
import numpy as np
import pytest

from aeroapps_b200 import mietables as mt


def _write_classic(path, raw):
    from scipy.io import netcdf_file
    with netcdf_file(path, "w") as nc:
        nlam, nrh, nbin = raw.bext.shape
        for name, n in (("lambda", nlam), ("rh", nrh), ("radius", nbin), ("nMom", raw.nmom), ("nPol", raw.npol)):
            nc.createDimension(name, n)
        def put(name, a, dims):                       # file order = Fortran order reversed
            v = nc.createVariable(name, "d", dims)
            v[:] = np.ascontiguousarray(np.asarray(a).T)
        put("lambda", raw.lam, ("lambda",)); put("rh", raw.rh, ("rh",)); put("rEff", raw.reff, ("radius", "rh"))
        for n, a in (("bext", raw.bext), ("bsca", raw.bsca), ("bbck", raw.bbck), ("g", raw.g), ("refreal", raw.refr), ("refimag", raw.refi)):
            put(n, a, ("radius", "rh", "lambda"))
        put("pback", raw.pback, ("nPol", "radius", "rh", "lambda"))
        put("pmom", raw.pmom, ("nPol", "nMom", "radius", "rh", "lambda"))
        put("growth_factor", raw.gf, ("radius", "rh")); put("rhop", raw.rhop, ("radius", "rh"))


def test_netcdf_classic_round_trip(tmp_path):
    raw = mt.synthetic_tables(nmom=5)["SS"]
    p = str(tmp_path / "optics_SS.nc")
    _write_classic(p, raw)
    got = mt.read_optics_file(p)
    for f in mt.RawTable.fields:
        a, b = getattr(raw, f), getattr(got, f)
        assert (a is None) == (b is None), f
        if a is not None:
            assert a.shape == b.shape and np.array_equal(a, b), f
    assert (got.nlam, got.nrh, got.nbin, got.nmom, got.npol) == (raw.nlam, raw.nrh, raw.nbin, raw.nmom, raw.npol)


def test_netcdf4_and_garbage_files_fail_with_a_reason(tmp_path):
    h5 = tmp_path / "optics_DU.v15_3.nc"
    h5.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    try:
        import netCDF4  # noqa: F401
        have = True
    except ImportError:
        try:
            import h5py  # noqa: F401
            have = True
        except ImportError:
            have = False
    if not have:
        with pytest.raises(RuntimeError, match="NetCDF-4/HDF5"):
            mt.read_optics_file(str(h5))
    junk = tmp_path / "optics.txt"
    junk.write_bytes(b"hello world, not a table")
    with pytest.raises(ValueError, match="not a NetCDF"):
        mt.read_optics_file(str(junk))


def test_getmiedims_keeps_the_reason(tmp_path):
    from aeroapps_b200 import MieObs_
    out = MieObs_.getmiedims(str(tmp_path / "missing.nc"))
    assert out[-1] == 1 and "missing.nc" in MieObs_.last_error
    raw = mt.synthetic_tables(nmom=3)["DU"]
    p = str(tmp_path / "optics_DU.nc")
    _write_classic(p, raw)
    assert MieObs_.getmiedims(p) == (raw.nlam, raw.nrh, raw.nbin, 3, raw.npol, 0)
