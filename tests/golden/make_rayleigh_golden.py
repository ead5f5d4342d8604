"""Generates tests/golden/rayleigh_python_reference.npz from the REFERENCE's own Python restatement of the Rayleigh spectral functions:
the methods `A` (cross section, cm^2 / molecule) and `F_air` (King factor) of the driver class in
src/Components/rtms/leo_vlidort/benchmark_polar_vlidort_hyperspectral.py:536-575, which restate the Fortran `A` / `F_air` of
VLIDORT_Rayleigh (RTSI/VLIDORT2p8p3/VLIDORT_ScatMod.F90:206-281) line by line.  The module itself cannot be imported (it needs the f2py
extension modules), so the two function definitions are cut out of its syntax tree and executed as they stand -- nothing is retyped.

Run in the build container only (reads /root/reference):   python tests/golden/make_rayleigh_golden.py
"""
import ast
import os
import types

import numpy as np

REF = "/root/reference/src/Components/rtms/leo_vlidort/benchmark_polar_vlidort_hyperspectral.py"


def reference_functions():
    tree = ast.parse(open(REF).read())
    fns = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef):
            for f in node.body:
                if isinstance(f, ast.FunctionDef) and f.name in ("A", "F_air"):
                    fns[f.name] = f
    assert set(fns) == {"A", "F_air"}
    mod = ast.Module(body=[fns["A"], fns["F_air"]], type_ignores=[])
    ns = {"np": np}
    exec(compile(mod, REF, "exec"), ns)
    holder = types.SimpleNamespace()
    holder.F_air = lambda XX2, C: ns["F_air"](holder, XX2, C)
    return (lambda X: ns["A"](holder, X)), holder.F_air


if __name__ == "__main__":
    A, F_air = reference_functions()
    lam_um = np.array([0.30, 0.34, 0.354, 0.388, 0.412, 0.44, 0.47, 0.55, 0.66, 0.67, 0.865, 0.87, 1.02, 1.24, 1.64, 2.13, 2.25])
    a = np.array([A(x) for x in lam_um])
    f = np.array([F_air(1.0 / (x * x), 360.0e-6) for x in lam_um])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rayleigh_python_reference.npz")
    np.savez(out, lam_um=lam_um, cross_section_cm2=a, king_factor=f)
    print(out, a[:3], f[:3])
