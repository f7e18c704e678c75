"""ctypes loader for oracle/_ref/librtlsim.so -- the cycle model translated mechanically from the mounted Verilog
(oracle/rtl2c/v2c.py) plus its test bench (oracle/rtl2c/tb.cpp).  TEST INFRASTRUCTURE ONLY: used by
tools/make_rtl_golden.py (which freezes its outputs into tests/golden/rtl_*.npz) and by tests/test_rtl_pin.py.

The library exists only where /root/reference is mounted (this container) or where the prebuilt .so travelled to
(the GPU box); `available()` says which.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "librtlsim.so")
_lib = None

EXT_FIELDS = ("score", "aw", "qle", "tle", "gtle", "gscore", "max_off")     # ap_return_0..6, sw_pe_array_sw_extend.v:117-123


def available() -> bool:
    return os.path.exists(_LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_LIB_PATH)
        _lib.rtl_sw_extend.restype = C.c_long
        _lib.rtl_sw_extend.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        _lib.rtl_pe_array.restype = C.c_long
        _lib.rtl_pe_array.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_long, C.c_uint64]
    return _lib


def sw_extend(query, target, h0, w, o_ins=6, e_ins=1, o_del=6, e_del=1, reg_score=0, max_ins=1000, max_del=1000,
              ret_in=(-1, -1, -1, -1, -1), scramble_seed=0):
    """One sw_extend call on the translated RTL.  Returns (7-tuple as int32 array in EXT_FIELDS order, clocks)."""
    bases = np.concatenate([np.asarray(query, dtype=np.uint8), np.asarray(target, dtype=np.uint8)])
    sc = np.array([o_ins, e_ins, o_del, e_del, w, h0, reg_score, max_ins, max_del], dtype=np.int32)
    ri = np.array(ret_in, dtype=np.int32)
    out = np.zeros(7, dtype=np.int32)
    clocks = lib().rtl_sw_extend(bases.ctypes.data, len(query), len(target), sc.ctypes.data, ri.ctypes.data,
                                 out.ctypes.data, int(scramble_seed))
    if clocks < 0:
        raise RuntimeError(f"rtl_sw_extend failed ({clocks})")
    return out, int(clocks)


def pe_array(tbb_words, rbb_fill=0xDEADBEEF, max_clocks=200_000_000, scramble_seed=0):
    """One batch through the translated sw_pe_array.  Returns (rbb image uint32[4096], result words written, clocks)."""
    tbb = np.ascontiguousarray(tbb_words, dtype=np.uint32)
    assert tbb.size == 65536
    rbb = np.full(4096, rbb_fill, dtype=np.uint32)
    n = C.c_int(0)
    clocks = lib().rtl_pe_array(tbb.ctypes.data, rbb.ctypes.data, C.byref(n), int(max_clocks), int(scramble_seed))
    if clocks < 0:
        raise RuntimeError(f"rtl_pe_array failed ({clocks})")
    return rbb, int(n.value), int(clocks)
