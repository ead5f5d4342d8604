"""ctypes loader for libaeroapps_b200.so (built in-tree by ``__graft_entry__.build()`` / csrc/Makefile).

Fails loudly: there is no fallback implementation of any entry point.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# AB200_LIB: developer override (instrumented builds made by tools/); the default is the in-tree library
LIB_PATH = os.environ.get("AB200_LIB") or os.path.join(_HERE, "libaeroapps_b200.so")


class LibraryMissing(RuntimeError):
    pass


_lib = None


def lib():
    """Return the loaded C-ABI library (ctypes.CDLL); raise LibraryMissing if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C aeroapps_b200/csrc` (there is no CPU fallback)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.ab200_last_error.restype = ctypes.c_char_p
        _lib.ab200_mie_get_table.restype = ctypes.c_long
    return _lib


def check(err, what):
    """Raise on a negative (library) error code; positive codes are the reference's own rc."""
    if err < 0:
        msg = lib().ab200_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed ({err}): {msg}")
    return err
