"""In-library multi-GPU: a host-pointer call with ab200_set_devices(2) splits its observations into one slab per
device (one host thread each, per-device contexts) and must return exactly what the single-device call returns -- results, Fourier counts,
MISSING rows and the FOCORR_NADIR state quirk across the slab boundary.  Needs two visible devices (gpurun --gpus 2); skipped otherwise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(built):
    from aeroapps_b200 import VLIDORT_POLAR_ as VP
    from aeroapps_b200._lib import lib
    from oracle import vlidort_glue as vg
    if lib().ab200_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    return VP, vg, lib()


@pytest.mark.parametrize("kind", ["lambert", "cx", "modis", "lambert_multigeom"])
def test_two_devices_equal_one(mods, kind):
    VP, vg, L = mods
    nobs, N, M = 23, 6, 1e15
    sc = vg.synthetic_scene(nobs, nch_nm=(470.0, 670.0), nmom=40, seed=321)
    sc["sza"][3] = 0.0                     # FOCORR_NADIR stays on from observation 3 on, also in the second device's slab
    sc["vza"][15] = M                      # a MISSING row in the second slab
    alpha = np.zeros_like(sc["ROT"]); flux = np.ones((2, nobs), order="F")

    def call():
        if kind == "lambert":
            out = VP.vector_lambert(sc["channels"], N, False, sc["ROT"], sc["depol"], alpha, sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"],
                                    sc["te"], sc["albedo"], sc["sza"], sc["raa"], sc["vza"], flux, M, 0)
        elif kind == "cx":
            out = VP.vector_cx(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"], sc["te"],
                               sc["u10m"], sc["v10m"], sc["mr"], sc["sza"], sc["raa"], sc["vza"], M, 0)
        elif kind == "modis":
            out = VP.vector_brdf_modis(sc["channels"], N, False, sc["ROT"], sc["depol"], sc["tau"], sc["ssa"], sc["pmom"], sc["pe"], sc["he"],
                                       sc["te"], sc["kernel_wt"], sc["param"], sc["sza"], sc["raa"], sc["vza"], M, 0)
        else:
            G = 4
            rng = np.random.default_rng(4)
            vza = np.asfortranarray(rng.uniform(0, 60, (nobs, G))); raa = np.asfortranarray(rng.uniform(0, 180, (nobs, G)))
            vza[15, 0] = M
            out = VP.vector_lambert_multigeom(sc["channels"], N, False, sc["ROT"], sc["depol"], alpha, sc["tau"], sc["ssa"], sc["pmom"], sc["pe"],
                                              sc["he"], sc["te"], sc["albedo"], sc["sza"], raa, vza, flux, M, 0)
        assert out[-1] == 0
        return out[:-1], VP.last_diag()

    assert L.ab200_set_devices(1) == 1
    one, d1 = call()
    assert L.ab200_set_devices(2) == 2
    try:
        two, d2 = call()
    finally:
        L.ab200_set_devices(1)
    for a, b in zip(one, two):
        assert np.array_equal(a, b)
    assert np.array_equal(d1[0], d2[0]) and np.array_equal(d1[1], d2[1])
    assert (one[0] == M).any() and (one[0] != M).any()
