"""ctypes loader for liblele_b200.so (the C-ABI CUDA library, include/lele_b200.h).

Fails loudly when the library is missing: there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "liblele_b200.so")


class LeleB200Error(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not os.path.exists(SO_PATH):
        raise LeleB200Error(
            f"{SO_PATH} is missing: build it with `python lele_b200/build.py` (nvcc, sm_100a). "
            "lele_b200 has no CPU fallback.")
    lib = C.CDLL(SO_PATH)
    lib.lele_b200_last_error.restype = C.c_char_p
    lib.lele_b200_launch_count.restype = C.c_ulonglong
    lib.lele_b200_launch_count.argtypes = [C.c_void_p]
    return lib


lib = _load()

vp = C.c_void_p
i32 = C.c_int
i64 = C.c_longlong
f32 = C.c_float
sz = C.c_size_t


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise LeleB200Error(f"{what}: {lib.lele_b200_last_error().decode(errors='replace')} (code {rc})")


def call(name: str, *args) -> None:
    fn = getattr(lib, name)
    check(fn(*args), name)
