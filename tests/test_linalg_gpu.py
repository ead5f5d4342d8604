"""GPU unit test of the CTA linear algebra behind K2b / K3 (aeroapps_b200/csrc/vl_linalg.cuh): the register-resident Gauss-Jordan
inversion with partial pivoting and the 16-byte-load matrix product, against numpy, for the problem sizes of the BASELINE configurations
(nstokes*nstreams = 24, 36, 48) and awkward ones (odd, small, needing row exchanges)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,threads", [(36, 128), (36, 256), (24, 128), (48, 256), (63, 384), (17, 128), (6, 128), (33, 128), (64, 512)])
def test_cta_inverse_and_product(built, n, threads):
    from aeroapps_b200._lib import lib, check
    L = lib()
    dp = C.POINTER(C.c_double)
    rng = np.random.default_rng(n * 1000 + threads)
    mats = [np.eye(n) + 0.3 * rng.standard_normal((n, n)),                       # general
            rng.standard_normal((n, n)),                                           # needs pivoting everywhere
            np.eye(n)[rng.permutation(n)] + 1e-3 * rng.standard_normal((n, n)),    # permutation-like: zero diagonal
            np.diag(10.0 ** rng.uniform(-6, 6, n)) @ (np.eye(n) + 0.1 * rng.standard_normal((n, n)))]   # badly row-scaled
    for A in mats:
        A = np.ascontiguousarray(A)
        E = np.zeros((n, n)); P = np.zeros((n, n))
        rc = check(L.ab200_selftest_linalg(n, threads, A.ctypes.data_as(dp), E.ctypes.data_as(dp), P.ctypes.data_as(dp)), "selftest_linalg")
        assert rc == 0
        Eref = np.linalg.inv(A)
        cond = np.linalg.cond(A)
        assert np.max(np.abs(E - Eref)) <= 1e-13 * cond * np.max(np.abs(Eref))
        assert np.max(np.abs(P - A @ E)) <= 1e-13 * np.max(np.abs(A)) * np.max(np.abs(E)) * n
        assert np.max(np.abs(A @ E - np.eye(n))) <= 1e-13 * cond
    # singular matrix -> reported, no crash
    A = np.ones((n, n)); E = np.zeros((n, n)); P = np.zeros((n, n))
    assert L.ab200_selftest_linalg(n, threads, A.ctypes.data_as(dp), E.ctypes.data_as(dp), P.ctypes.data_as(dp)) == 1
