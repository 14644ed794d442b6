"""ctypes binding of libhrp_b200.so (include/hrp.h).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libhrp_b200.so"
_lib = None


class HrpError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("kind", "B", "Hin", "Win", "Cin", "Cout", "kh", "kw", "stride", "pad", "relu")]


class ConvEpilogue(C.Structure):
    _fields_ = [
        ("scale", C.c_void_p), ("bias", C.c_void_p),
        ("pre", C.c_void_p * 3), ("up", C.c_void_p * 3), ("up_shift", C.c_int32 * 3),
        ("post", C.c_void_p), ("out", C.c_void_p), ("pool_out", C.c_void_p),
    ]


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise HrpError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the HoRoPose B200 path)")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.hrp_last_error.restype = C.c_char_p
        _lib.hrp_version.restype = C.c_char_p
        _lib.hrp_launch_count.restype = C.c_int64
        for name in dir(_lib):
            pass
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().hrp_last_error()
        raise HrpError(f"hrp error {rc}: {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(lib().hrp_launch_count())
