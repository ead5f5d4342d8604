"""aeroapps_b200 -- B200-native (sm_100a CUDA) drop-in for the AeroApps Mie-AOP + VLIDORT hot path.

The package holds only what that path needs: ``csrc/`` (CUDA kernels + the C ABI declared in
``include/aeroapps_b200.h``), and host-side mirrors of the reference's f2py modules
(``MieObs_``, ``VLIDORT_POLAR_``).  There is no CPU fallback: importing a mirror module without the
built CUDA library raises.
"""
from ._lib import lib, LibraryMissing, check  # noqa: F401

__all__ = ["lib", "LibraryMissing", "check"]
