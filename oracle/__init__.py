"""CPU oracle loader -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It wraps oracle/libbswref.so (plain-C restatement of the reference's
sw_extend / proc_element path, see ksw_extend_ref.c) with numpy-friendly helpers.

Parity status: "parity unpinned" by the reference (RTL only, no golden vectors, no simulator);
pinned by KATs + oracle/matrix_model.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbswref.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ksw_extend_ref.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class Params(C.Structure):
    _fields_ = [("mat", C.c_int8 * 25), ("o_del", C.c_int32), ("e_del", C.c_int32),
                ("o_ins", C.c_int32), ("e_ins", C.c_int32), ("zdrop", C.c_int32),
                ("end_bonus", C.c_int32)]


class Params2(C.Structure):
    _fields_ = [("p", Params), ("w", C.c_int32), ("pen_clip5", C.c_int32), ("pen_clip3", C.c_int32)]


class SeedTask(C.Structure):
    _fields_ = [("q_left", C.c_void_p), ("q_right", C.c_void_p), ("t_left", C.c_void_p),
                ("t_right", C.c_void_p), ("qlen", C.c_int32 * 2), ("tlen", C.c_int32 * 2),
                ("init_score", C.c_int32), ("qbeg", C.c_int32), ("h0", C.c_int32), ("id", C.c_uint32)]


RESULT_DTYPE = np.dtype([("score", "<i4"), ("qle", "<i4"), ("tle", "<i4"),
                         ("gtle", "<i4"), ("gscore", "<i4"), ("max_off", "<i4")])
ALN_DTYPE = np.dtype([("id", "<u4"), ("qb", "<i4"), ("qe", "<i4"), ("rb", "<i4"), ("re", "<i4"),
                      ("score", "<i4"), ("truesc", "<i4"), ("w", "<i4")])

_lib = None


def use_library(path: str) -> None:
    """Load the oracle from another build of ksw_extend_ref.c (bench.py: the -march=native build made on the box)."""
    global _lib, _LIB_PATH
    _LIB_PATH = path
    _lib = None
    lib()


def lib():
    global _lib
    if _lib is None:
        if _LIB_PATH == os.path.join(_HERE, "libbswref.so"):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.bswref_extend.restype = C.c_int
        _lib.bswref_extend.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int64)]
        _lib.bswref_extend_batch.restype = None
        _lib.bswref_extend_batch.argtypes = [C.POINTER(Params), C.c_int, C.c_int64] + [C.c_void_p] * 8 + [C.c_int]
        _lib.bswref_chain2aln_batch.restype = None
        _lib.bswref_chain2aln_batch.argtypes = [C.POINTER(Params2), C.c_int, C.c_int64, C.c_void_p,
                                                C.c_void_p, C.POINTER(C.c_int64), C.c_int]
        _lib.bswref_sw_extend_rtl.restype = None
        _lib.bswref_sw_extend_rtl.argtypes = [C.POINTER(Params), C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 5 + \
                                             [C.c_void_p, C.POINTER(C.c_int64)]
        _lib.bswref_chain2aln_rtl.restype = None
        _lib.bswref_chain2aln_rtl.argtypes = [C.POINTER(Params2), C.POINTER(SeedTask), C.c_void_p, C.c_void_p,
                                              C.POINTER(C.c_int64)]
        _lib.bswref_sw_extend_rtl8.restype = None
        _lib.bswref_sw_extend_rtl8.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 9 + [C.c_void_p]
        _lib.bswref_global.restype = C.c_int
        _lib.bswref_global.argtypes = [C.POINTER(Params), C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        _lib.bswref_max_threads.restype = C.c_int
        _lib.bswref_clamp_w.restype = C.c_int
        _lib.bswref_clamp_w.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int]
    return _lib


def make_params(mat=None, o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=100, end_bonus=5, a=1, b=4) -> Params:
    p = Params()
    if mat is None:
        mat = bwa_fill_scmat(a, b)
    mat = np.asarray(mat, dtype=np.int8).reshape(25)
    for i in range(25):
        p.mat[i] = int(mat[i])
    p.o_del, p.e_del, p.o_ins, p.e_ins, p.zdrop, p.end_bonus = o_del, e_del, o_ins, e_ins, zdrop, end_bonus
    return p


def bwa_fill_scmat(a: int = 1, b: int = 4) -> np.ndarray:
    """BWA's 5x5 matrix: +a on the diagonal, -b off it, -1 for anything involving N
    (the constants hard-wired in sw_pe_array_sw_extend.v:1915-1940 for a=1,b=4)."""
    m = np.full((5, 5), -b, dtype=np.int8)
    for i in range(4):
        m[i, i] = a
    m[4, :] = -1
    m[:, 4] = -1
    return m.reshape(25)


def max_threads() -> int:
    return int(lib().bswref_max_threads())


def extend_one(params: Params, query, target, h0: int, w: int, variant: int = 1):
    q = np.ascontiguousarray(query, dtype=np.uint8)
    t = np.ascontiguousarray(target, dtype=np.uint8)
    out = np.zeros(1, dtype=RESULT_DTYPE)
    cells = C.c_int64(0)
    lib().bswref_extend(C.byref(params), variant, len(q), q.ctypes.data, len(t), t.ctypes.data,
                        int(w), int(h0), out.ctypes.data, C.byref(cells))
    return out[0], int(cells.value)


def sw_extend_rtl(params: Params, query, target, h0: int, w: int, reg_score: int, max_ins: int, max_del: int):
    """One whole RTL sw_extend invocation in int32: host-supplied band clamp, two band tries with the maxima carried
    from try to try, no z-drop.  Returns (int32[7] = score, aw, qle, tle, gtle, gscore, max_off; cells)."""
    q = np.ascontiguousarray(query, dtype=np.uint8)
    t = np.ascontiguousarray(target, dtype=np.uint8)
    out = np.zeros(7, dtype=np.int32)
    cells = C.c_int64(0)
    lib().bswref_sw_extend_rtl(C.byref(params), len(q), q.ctypes.data, len(t), t.ctypes.data, int(w), int(h0),
                               int(reg_score), int(max_ins), int(max_del), out.ctypes.data, C.byref(cells))
    return out, int(cells.value)


def sw_extend_rtl8(query, target, h0, w, o_ins=6, e_ins=1, o_del=6, e_del=1, reg_score=0, max_ins=1000, max_del=1000):
    """One sw_extend invocation at the RTL's widths (oracle/rtl_width_model.c): what the FPGA returns, wraps included.
    int32[7] = score, aw, qle, tle, gtle, gscore, max_off.  Claimed for qlen <= 127 (see the file header)."""
    q = np.ascontiguousarray(query, dtype=np.uint8)
    t = np.ascontiguousarray(target, dtype=np.uint8)
    out = np.zeros(7, dtype=np.int32)
    lib().bswref_sw_extend_rtl8(len(q), q.ctypes.data, len(t), t.ctypes.data, int(o_ins), int(e_ins), int(o_del), int(e_del),
                                int(w), int(h0), int(reg_score), int(max_ins), int(max_del), out.ctypes.data)
    return out


def rtl_envelope(qlen, tlen, h0, w, o_del=6, e_del=1, o_ins=6, e_ins=1) -> bool:
    """True when the FPGA's 8-bit datapath computes one sw_extend call without wrapping, i.e. when the RTL and int32
    ksw_extend2 must agree (checked against the translated RTL on tests/golden/rtl_sw_extend*.npz:
    no task inside this envelope differs, tasks outside it do).
      scores      h0 + qlen*max(mat) <= 127: H/E/F/max are 8 bit, compared signed (sw_pe_array_sw_extend.v:155-159,1943-1945)
      columns     qlen <= 127: mj is sign-extended in max_off and in the narrowing compares (sx:1654,1336,1547)
      first col.  h0 - o_del - e_del*tlen >= -128: h1 is a running 8-bit subtraction clipped by its sign bit (sx:1795,890-907)
      first row   h0 - o_ins - e_ins*qlen >= -128: same for the row-0 generator (sx:1072,1975)
      band        w <= 63: w << 1 is a signed 8-bit value in the second band try (sx:770,775)"""
    return bool(qlen <= 127 and h0 + qlen <= 127 and h0 - o_del - e_del * tlen >= -128
                and h0 - o_ins - e_ins * qlen >= -128 and 0 < w <= 63 and 0 < h0 and 0 < qlen and 0 < tlen <= 2047)


def extend_batch(params: Params, qbuf, qoff, tbuf, toff, h0, w, variant: int = 1, nthreads: int = 0):
    """Level-1 oracle over a flat batch.  Returns (results[RESULT_DTYPE], cells_per_task[int64])."""
    qbuf = np.ascontiguousarray(qbuf, dtype=np.uint8)
    tbuf = np.ascontiguousarray(tbuf, dtype=np.uint8)
    qoff = np.ascontiguousarray(qoff, dtype=np.int64)
    toff = np.ascontiguousarray(toff, dtype=np.int64)
    h0 = np.ascontiguousarray(h0, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.int32)
    n = len(h0)
    assert len(qoff) == n + 1 and len(toff) == n + 1 and len(w) == n
    out = np.zeros(n, dtype=RESULT_DTYPE)
    cells = np.zeros(n, dtype=np.int64)
    lib().bswref_extend_batch(C.byref(params), variant, n, qbuf.ctypes.data, qoff.ctypes.data,
                              tbuf.ctypes.data, toff.ctypes.data, h0.ctypes.data, w.ctypes.data,
                              out.ctypes.data, cells.ctypes.data, int(nthreads))
    return out, cells


def chain2aln_batch(params2: Params2, seed_tasks, variant: int = 1, nthreads: int = 0):
    """Level-2 oracle.  seed_tasks: ctypes array of SeedTask.  Returns (records[ALN_DTYPE], cells_total)."""
    n = len(seed_tasks)
    out = np.zeros(n, dtype=ALN_DTYPE)
    cells = C.c_int64(0)
    lib().bswref_chain2aln_batch(C.byref(params2), variant, n, C.addressof(seed_tasks) if n else None,
                                 out.ctypes.data, C.byref(cells), int(nthreads))
    return out, int(cells.value)


def chain2aln_rtl(params2: Params2, seed_tasks, gaps):
    """Level 2 exactly as the RTL sequences it: gaps[i] = (max_ins_left, max_del_left, max_ins_right, max_del_right)
    from the TBB (host-supplied band clamp), each side one whole sw_extend invocation (sw_extend_rtl)."""
    n = len(seed_tasks)
    out = np.zeros(n, dtype=ALN_DTYPE)
    gaps = np.ascontiguousarray(gaps, dtype=np.int32).reshape(n, 4)
    cells = C.c_int64(0)
    for i in range(n):
        lib().bswref_chain2aln_rtl(C.byref(params2), C.byref(seed_tasks[i]), gaps[i].ctypes.data,
                                   out[i:i + 1].ctypes.data, C.byref(cells))
    return out, int(cells.value)


def global_align(params: Params, query, target, w: int, max_cigar: int = 512):
    """ksw_global2 restated (banded global alignment + traceback; SURVEY 8 f.4, parity unpinned -- see the C file).
    Returns (score, cigar uint32[n] in BAM encoding len << 4 | op with 0 = M, 1 = I, 2 = D), or (score, None) when the
    alignment needs more than max_cigar operations."""
    q = np.ascontiguousarray(query, dtype=np.uint8)
    t = np.ascontiguousarray(target, dtype=np.uint8)
    cig = np.zeros(max_cigar, dtype=np.uint32)
    n = C.c_int(0)
    score = lib().bswref_global(C.byref(params), len(q), q.ctypes.data, len(t), t.ctypes.data, int(w), max_cigar, cig.ctypes.data, C.byref(n))
    return int(score), (cig[:n.value].copy() if n.value >= 0 else None)
