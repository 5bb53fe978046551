"""ctypes loader for libgroma_b200.so.

Signatures are parsed from include/groma_b200.h, so the header is the single source of truth for the C ABI.
There is no CPU fallback: if the library is missing or a kernel returns an error the call raises.
"""
from __future__ import annotations

import ctypes
import re
from functools import lru_cache
from pathlib import Path

ROOT = Path(__file__).resolve().parent
HEADER = ROOT.parent / "include" / "groma_b200.h"
import os as _os
LIB_PATH = Path(_os.environ.get("GROMA_B200_LIB", str(ROOT / "lib" / "libgroma_b200.so")))   # override = A/B tuning builds only

_CTYPES = {
    "int32_t": ctypes.c_int32,
    "int64_t": ctypes.c_int64,
    "float": ctypes.c_float,
}

ERRORS = {1: "GROMA_ERR_ARG", 2: "GROMA_ERR_ALIGN", 3: "GROMA_ERR_CUDA", 4: "GROMA_ERR_DRIVER",
          5: "GROMA_ERR_TMA_ENCODE", 6: "GROMA_ERR_UNSUPPORTED"}


class GromaError(RuntimeError):
    pass


def parse_header(path: Path = HEADER) -> dict[str, list[tuple[str, str]]]:
    """Return {function name: [(c_type, arg_name), ...]} for every `int32_t groma_*(...)` declaration."""
    text = path.read_text()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    out: dict[str, list[tuple[str, str]]] = {}
    for m in re.finditer(r"int32_t\s+(groma_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2)
        parsed = []
        for a in args.split(","):
            a = " ".join(a.split())
            if not a:
                continue
            mm = re.match(r"(.*?)(\w+)$", a)
            ctype, an = mm.group(1).strip(), mm.group(2)
            parsed.append((ctype, an))
        out[name] = parsed
    return out


def _to_ctype(ctype: str):
    if "*" in ctype:
        return ctypes.c_void_p
    base = ctype.replace("const", "").strip()
    return _CTYPES[base]


@lru_cache(maxsize=1)
def load() -> ctypes.CDLL:
    if not LIB_PATH.exists():
        raise GromaError(f"{LIB_PATH} not built -- run `python -m groma_b200.build` (no CPU fallback exists)")
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, args in parse_header().items():
        fn = getattr(lib, name)  # raises AttributeError if the symbol is not exported
        fn.restype = ctypes.c_int32
        fn.argtypes = [_to_ctype(t) for t, _ in args]
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise GromaError(f"{what} failed: {ERRORS.get(rc, rc)}")
