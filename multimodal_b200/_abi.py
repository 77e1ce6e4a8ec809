"""ctypes prototypes for every symbol declared in include/mmb200.h."""
from __future__ import annotations

import ctypes as C

vp, ll, i32, f32 = C.c_void_p, C.c_longlong, C.c_int, C.c_float

PROTOTYPES = {
    "mmb_version": (i32, []),
    "mmb_gemm_bf16": (i32, [vp, ll, i32, vp, ll, i32, vp, ll, vp, ll, i32, i32, i32, i32, i32, f32, vp, vp, ll, i32, i32, vp]),
}


def declare(lib: C.CDLL) -> None:
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
