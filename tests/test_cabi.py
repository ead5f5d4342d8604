"""The drop-in boundary from plain C: tests/cabi/test_cabi.c includes include/aeroapps_b200.h as C99, links against
libaeroapps_b200.so alone and drives ROT_CALC -> VECTOR_CX -> SCALAR_LAMBERT with malloc'ed Fortran-ordered buffers.

CPU part: the headers are valid C (gcc -std=c99 -pedantic -Werror) and the harness links (every symbol it binds resolves).
GPU part: the harness runs on the box; its outputs equal the ctypes path bit for bit and the oracle to the parity bar."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cabi", "test_cabi.c")
BUILD = os.path.join(ROOT, "tests", "cabi", "_build")
EXE = os.path.join(BUILD, "test_cabi")


def _compile():
    from aeroapps_b200 import _lib
    os.makedirs(BUILD, exist_ok=True)
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
           "-L", libdir, "-laeroapps_b200", f"-Wl,-rpath,{libdir}", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_headers_are_c99_and_harness_links(built):
    for h in ("aeroapps_b200.h",):                      # it includes aeroapps_b200_rt.h
        r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", "-include",
                            os.path.join(ROOT, "include", "aeroapps_b200.h"), "/dev/null"], capture_output=True, text=True)
        assert r.returncode == 0, (h, r.stderr)
    assert os.path.exists(_compile())


@pytest.mark.gpu
def test_c_harness_matches_ctypes_and_oracle(built, tmp_path):
    from aeroapps_b200 import VLIDORT_POLAR_ as VP
    from oracle import vlidort_glue as vg
    exe = _compile()
    km, nobs, N, nmom = 72, 5, 6, 32
    sc = vg.synthetic_scene(nobs, nch_nm=(470.0, 870.0), nmom=nmom, seed=77)
    nch = 2
    scene, out = tmp_path / "scene.bin", tmp_path / "out.bin"
    F = lambda a: np.asfortranarray(a, dtype=np.float64).ravel(order="F")
    with open(scene, "wb") as f:
        np.array([km, nch, nobs, N, 0, nmom, sc["pmom"].shape[4]], np.int32).tofile(f)
        for k in ("channels", "pe", "he", "te", "tau", "ssa", "g", "pmom", "albedo", "u10m", "v10m", "mr", "sza", "raa", "vza"):
            F(sc[k]).tofile(f)
    r = subprocess.run([exe, str(scene), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(out, "rb").read()
    nd = km * nobs * nch + nch + 9 * nobs * nch
    d = np.frombuffer(raw, np.float64, nd); ints = np.frombuffer(raw, np.int32, offset=8 * nd)
    o = 0
    def take(shape):
        nonlocal o
        n = int(np.prod(shape)); a = d[o:o + n].reshape(shape, order="F"); o += n
        return a
    ROT, depol = take((km, nobs, nch)), take((nch,))
    cx = [take((nobs, nch)) for _ in range(7)]
    sl = [take((nobs, nch)) for _ in range(2)]
    assert list(ints[:3]) == [0, 0, 0]
    nf_c, st_c = ints[3:3 + nobs * nch], ints[3 + nobs * nch:3 + 2 * nobs * nch]
    assert (st_c == 0).all()
    # the same calls through the ctypes mirror: bit-identical
    ROT2, depol2, rc = VP.rot_calc(sc["channels"], sc["pe"], sc["he"], sc["te"], 1e15, 0)
    assert rc == 0 and np.array_equal(ROT, ROT2) and np.array_equal(depol, depol2)
    ref = VP.vector_cx(sc["channels"], N, False, ROT2, depol2, sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"], sc["te"], sc["u10m"],
                       sc["v10m"], sc["mr"], sc["sza"], sc["raa"], sc["vza"], 1e15, 0)
    for a, b in zip(cx, ref[:7]):
        assert np.array_equal(a, b)
    ref_nf_cx = VP.last_diag()[0].ravel()
    ref = VP.scalar_lambert(sc["channels"], N, False, ROT2, depol2, sc["tau"], sc["ssa"], sc["g"], sc["pmom"], sc["pe"], sc["he"], sc["te"],
                            sc["albedo"], sc["sza"], sc["raa"], sc["vza"], 1e15, 0)
    assert np.array_equal(sl[0], ref[0]) and np.array_equal(sl[1], ref[1])
    nf, st = VP.last_diag()
    assert np.array_equal(nf.ravel(), nf_c)
    # and the oracle, parity bar 1e-8
    orc = vg.vlidort_run_surface("cx", 3, N, False, ROT2, depol2, sc["tau"], sc["ssa"], sc["pmom"], sc["he"], sc["sza"], sc["raa"], sc["vza"],
                                 u10m=sc["u10m"], v10m=sc["v10m"], mr=sc["mr"])
    assert np.max(np.abs(cx[0] - orc["radiance"]) / np.abs(orc["radiance"])) < 1e-8
    assert np.array_equal(ref_nf_cx, orc["nfourier"].ravel())
