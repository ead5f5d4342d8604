"""The C-ABI library loads and exports every function the public headers declare (no compute calls)."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(ab200_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_all_declared_symbols(built):
    from aeroapps_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 10
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    assert lib.ab200_version() >= 100


def test_product_does_not_import_oracle():
    for path in glob.glob(os.path.join(ROOT, "aeroapps_b200", "**", "*"), recursive=True):
        if path.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
            txt = open(path).read()
            assert "oracle" not in txt.replace("the oracle", "").lower() or "import oracle" not in txt, path
            assert "from oracle" not in txt and "import oracle" not in txt and "mie_oracle" not in txt, path
