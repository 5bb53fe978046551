"""CPU-side checks of the C-ABI boundary: the library builds, loads, and exports every symbol the header declares."""
import ctypes

from groma_b200 import lib as L


def test_header_parses_and_is_nonempty():
    decls = L.parse_header()
    assert len(decls) >= 25
    for name, args in decls.items():
        assert name.startswith("groma_")
        assert args[-1][1] == "stream", name


def test_library_exports_every_declared_symbol():
    from groma_b200.build import build
    path = build(verbose=False)
    lib = ctypes.CDLL(str(path))
    for name in L.parse_header():
        assert hasattr(lib, name), f"{name} declared in include/groma_b200.h but not exported"
    L.load()  # sets argtypes from the header without touching a GPU
