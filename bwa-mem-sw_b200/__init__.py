"""bsw_b200 -- host-side mirror of the reference's seed-extension interface over libbsw.so (C ABI: include/bsw.h).

The reference (peterpengwei/bwa-mem-sw) exposes this path as RTL modules; the three call levels here keep their
names and argument meaning:

  * ``Context.sw_extend_batch``      <-> ``sw_pe_array_sw_extend``   (one ksw_extend2 call per task;
                                         ports sw_pe_array_sw_extend.v:96-123, returns in the order of :117-123)
  * ``Context.proc_element_batch``   <-> ``sw_pe_array_proc_element`` (left + right extension, band retry, clip;
                                         sw_pe_array_proc_element.v:1593-1685, record of :1187-1205)
  * ``Context.pe_array_batch``       <-> ``sw_pe_array`` fed by ``batch_manager``/``task_parse``/``fill_resulBuf``
                                         (TBB image in, RBB image out; tbb.v:163-194, rbb.v:117-167)

This module is plumbing only (ctypes + numpy).  All compute happens in the CUDA kernels inside libbsw.so; there is
no CPU fallback -- ``Context()`` raises if the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

# libbsw.so's host pipeline keeps ~100 streams busy; more hardware queues than the default 8 avoids false dependencies
# between them.  The driver reads this when the CUDA context is created, so set it before anything touches the GPU.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libbsw.so")
SYNTH_PATH = os.path.join(_HERE, "libbsw_synth.so")
EMU_PATH = os.path.join(_HERE, "libbsw_emu.so")

BSW_OK, BSW_EINVAL, BSW_ECUDA, BSW_ENOMEM, BSW_ERANGE, BSW_EWIRE, BSW_EBUSY = 0, -1, -2, -3, -4, -5, -6
ERR_NAMES = {0: "BSW_OK", -1: "BSW_EINVAL", -2: "BSW_ECUDA", -3: "BSW_ENOMEM", -4: "BSW_ERANGE", -5: "BSW_EWIRE", -6: "BSW_EBUSY"}
TBB_WORDS, RBB_WORDS = 65536, 4096

RESULT_DTYPE = np.dtype([("score", "<i4"), ("qle", "<i4"), ("tle", "<i4"),
                         ("gtle", "<i4"), ("gscore", "<i4"), ("max_off", "<i4")])
ALN_DTYPE = np.dtype([("id", "<u4"), ("qb", "<i4"), ("qe", "<i4"), ("rb", "<i4"), ("re", "<i4"),
                      ("score", "<i4"), ("truesc", "<i4"), ("w", "<i4")])


class BswError(RuntimeError):
    def __init__(self, code: int, text: str = ""):
        self.code = code
        super().__init__(f"{ERR_NAMES.get(code, code)}: {text}")


class Params(C.Structure):
    _fields_ = [("mat", C.c_int8 * 25), ("o_del", C.c_int32), ("e_del", C.c_int32),
                ("o_ins", C.c_int32), ("e_ins", C.c_int32), ("zdrop", C.c_int32), ("end_bonus", C.c_int32)]


class Params2(C.Structure):
    _fields_ = [("p", Params), ("w", C.c_int32), ("pen_clip5", C.c_int32), ("pen_clip3", C.c_int32)]


class Task(C.Structure):
    _fields_ = [("query", C.c_void_p), ("target", C.c_void_p), ("qlen", C.c_int32), ("tlen", C.c_int32),
                ("h0", C.c_int32), ("w", C.c_int32)]


class GlobalTask(C.Structure):
    _fields_ = [("query", C.c_void_p), ("target", C.c_void_p), ("qlen", C.c_int32), ("tlen", C.c_int32), ("w", C.c_int32)]


class SeedTask(C.Structure):
    _fields_ = [("q_left", C.c_void_p), ("q_right", C.c_void_p), ("t_left", C.c_void_p), ("t_right", C.c_void_p),
                ("qlen", C.c_int32 * 2), ("tlen", C.c_int32 * 2),
                ("init_score", C.c_int32), ("qbeg", C.c_int32), ("h0", C.c_int32), ("id", C.c_uint32)]


class Ticket(C.Structure):
    _fields_ = [("slot", C.c_int32), ("seq", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("tasks", C.c_uint64), ("cells_band", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("kernel_ms", C.c_double), ("pack_ms", C.c_double), ("wall_ms", C.c_double)]


class IntPeak(C.Structure):
    _fields_ = [("iadd_tops", C.c_double), ("vimnmx_tops", C.c_double), ("dpx_tops", C.c_double),
                ("mix_tops", C.c_double), ("dual_tops", C.c_double), ("sm_clock_mhz", C.c_double), ("sm_count", C.c_int)]


class SynthCfg(C.Structure):
    _fields_ = [("read_len_min", C.c_int32), ("read_len_max", C.c_int32), ("seed_min", C.c_int32), ("seed_max", C.c_int32),
                ("long_mode", C.c_int32), ("qlen_min", C.c_int32), ("qlen_max", C.c_int32), ("h0_min", C.c_int32),
                ("h0_max", C.c_int32), ("w", C.c_int32), ("a", C.c_int32), ("o", C.c_int32),
                ("sub", C.c_double), ("ins", C.c_double), ("del_", C.c_double), ("unrelated_frac", C.c_double),
                ("n_frac", C.c_double), ("seed", C.c_uint64)]


def build(force: bool = False, verbose: bool = False) -> None:
    """Compile libbsw.so / libbsw_synth.so / libbsw_emu.so in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    cmd = ["make", "-C", _CSRC, "-j8"] + (["-B"] if force else [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("building libbsw.so failed")


def bwa_fill_scmat(a: int = 1, b: int = 4) -> np.ndarray:
    """BWA's 5x5 matrix (+a diagonal, -b off-diagonal, -1 with N): the constants of sw_pe_array_sw_extend.v:1915-1940."""
    m = np.full((5, 5), -b, dtype=np.int8)
    for i in range(4):
        m[i, i] = a
    m[4, :] = -1
    m[:, 4] = -1
    return m.reshape(25)


def make_params(mat=None, o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=100, end_bonus=5, a=1, b=4) -> Params:
    p = Params()
    mat = bwa_fill_scmat(a, b) if mat is None else np.asarray(mat, dtype=np.int8).reshape(25)
    for i in range(25):
        p.mat[i] = int(mat[i])
    p.o_del, p.e_del, p.o_ins, p.e_ins, p.zdrop, p.end_bonus = o_del, e_del, o_ins, e_ins, zdrop, end_bonus
    return p


def make_params2(params: Params | None = None, w=100, pen_clip5=5, pen_clip3=5, **kw) -> Params2:
    P = Params2()
    src = params if params is not None else make_params(**kw)
    C.memmove(C.byref(P.p), C.byref(src), C.sizeof(Params))
    P.w, P.pen_clip5, P.pen_clip3 = w, pen_clip5, pen_clip3
    return P


_lib = None


def lib() -> C.CDLL:
    """Load libbsw.so.  Fails loudly when it has not been built -- there is no other implementation to fall back to."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (nvcc) first; there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
        L.bsw_init.argtypes = [C.POINTER(vp), vp, i32, i32]
        L.bsw_destroy.argtypes = [vp]; L.bsw_destroy.restype = None
        L.bsw_last_error.argtypes = [vp]; L.bsw_last_error.restype = C.c_char_p
        L.bsw_version.restype = C.c_char_p
        L.bsw_set_option.argtypes = [vp, C.c_char_p, i64]
        L.bsw_num_devices.argtypes = [vp]
        L.bsw_extend_batch.argtypes = [vp, C.POINTER(Params), vp, C.c_size_t, vp]
        L.bsw_extend_batch_flat.argtypes = [vp, C.POINTER(Params), vp, vp, vp, vp, vp, vp, C.c_size_t, vp, vp]
        L.bsw_chain2aln_batch.argtypes = [vp, C.POINTER(Params2), vp, C.c_size_t, vp]
        L.bsw_global_batch.argtypes = [vp, C.POINTER(Params), vp, C.c_size_t, i32, vp, vp, vp]
        L.bsw_fpga_batch.argtypes = [vp, vp, vp, C.POINTER(i32)]
        L.bsw_fpga_envelope.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.bsw_tbb_encode.argtypes = [C.POINTER(Params2), vp, C.c_size_t, vp]
        L.bsw_rbb_decode.argtypes = [vp, C.c_size_t, vp]
        L.bsw_submit.argtypes = [vp, C.POINTER(Params), vp, C.c_size_t, vp, C.POINTER(Ticket)]
        L.bsw_poll.argtypes = [vp, C.POINTER(Ticket)]
        L.bsw_wait.argtypes = [vp, C.POINTER(Ticket)]
        L.bsw_resident_create.argtypes = [vp, C.POINTER(Params), vp, vp, vp, vp, vp, vp, C.c_size_t, C.POINTER(vp)]
        L.bsw_resident_run.argtypes = [vp, vp, C.POINTER(C.c_double), C.POINTER(u64), C.POINTER(u64)]
        L.bsw_resident_fetch.argtypes = [vp, vp, vp, vp]
        L.bsw_resident_free.argtypes = [vp, vp]; L.bsw_resident_free.restype = None
        L.bsw_host_register.argtypes = [vp, vp, C.c_size_t]
        L.bsw_host_unregister.argtypes = [vp, vp]
        L.bsw_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.bsw_reset_stats.argtypes = [vp]
        L.bsw_measure_int_peak.argtypes = [vp, i32, C.POINTER(IntPeak)]
        _lib = L
    return _lib


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _flat_args(qbuf, qoff, tbuf, toff, h0, w):
    qbuf, tbuf = _u8(qbuf), _u8(tbuf)
    qoff = np.ascontiguousarray(qoff, dtype=np.int64)
    toff = np.ascontiguousarray(toff, dtype=np.int64)
    n = len(qoff) - 1
    h0 = np.ascontiguousarray(np.broadcast_to(np.asarray(h0, dtype=np.int32), (n,)))
    w = np.ascontiguousarray(np.broadcast_to(np.asarray(w, dtype=np.int32), (n,)))
    return qbuf, qoff, tbuf, toff, h0, w, n


class Resident:
    """A batch whose packed inputs live in HBM (measurement of the kernels alone)."""

    def __init__(self, ctx: "Context", handle):
        self.ctx, self.handle = ctx, handle

    def run(self):
        ms, cells, nl = C.c_double(0), C.c_uint64(0), C.c_uint64(0)
        self.ctx._check(lib().bsw_resident_run(self.ctx.handle, self.handle, C.byref(ms), C.byref(cells), C.byref(nl)))
        return ms.value, int(cells.value), int(nl.value)

    def fetch(self, n: int):
        out = np.zeros(n, dtype=RESULT_DTYPE)
        cells = np.zeros(n, dtype=np.uint32)
        self.ctx._check(lib().bsw_resident_fetch(self.ctx.handle, self.handle, out.ctypes.data, cells.ctypes.data))
        return out, cells

    def free(self):
        if self.handle:
            lib().bsw_resident_free(self.ctx.handle, self.handle)
            self.handle = None


class Context:
    """bsw_ctx: one context drives one or several GPUs of the box (host-sharded, no collective)."""

    def __init__(self, devices=None, streams_per_device: int = 2, **options):
        L = lib()
        self.handle = C.c_void_p()
        if devices is None:
            rc = L.bsw_init(C.byref(self.handle), None, 0, streams_per_device)
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = L.bsw_init(C.byref(self.handle), arr, len(devices), streams_per_device)
        if rc != BSW_OK:
            raise BswError(rc, "bsw_init failed (no usable sm_100 CUDA device? there is no CPU fallback)")
        for k, v in options.items():
            self.set_option(k, v)
        for kv in filter(None, os.environ.get("BSW_OPTIONS", "").split(",")):      # experiments: BSW_OPTIONS="raw_inputs=3,slots=2"
            k, _, v = kv.partition("=")
            self.set_option(k.strip(), int(v))

    def close(self):
        if self.handle:
            lib().bsw_destroy(self.handle)
            self.handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != BSW_OK:
            raise BswError(rc, (lib().bsw_last_error(self.handle) or b"").decode())

    def set_option(self, key: str, value: int):
        self._check(lib().bsw_set_option(self.handle, key.encode(), int(value)))

    @property
    def num_devices(self) -> int:
        return lib().bsw_num_devices(self.handle)

    # ---- level 1: sw_extend ----
    def register_host(self, arr: np.ndarray) -> None:
        """Page-lock a long-lived buffer in place (bsw_host_register): flat batches whose bases live in registered uint8
        buffers skip the host staging pass, and a registered result array receives its records straight from the device.
        Keep the array alive until unregister_host / close."""
        assert arr.flags.c_contiguous
        self._check(lib().bsw_host_register(self.handle, arr.ctypes.data, arr.nbytes))

    def unregister_host(self, arr: np.ndarray) -> None:
        self._check(lib().bsw_host_unregister(self.handle, arr.ctypes.data))

    def sw_extend_batch(self, params: Params, qbuf, qoff, tbuf, toff, h0, w, want_cells: bool = True, out=None, cells=None):
        """Flat layout: task i's query is qbuf[qoff[i]:qoff[i+1]].  Returns (results[RESULT_DTYPE], cells[uint32]).
        out / cells: optional caller-owned result arrays (the C ABI writes into caller buffers; reuse them across calls)."""
        qbuf, qoff, tbuf, toff, h0, w, n = _flat_args(qbuf, qoff, tbuf, toff, h0, w)
        if out is None:
            out = np.zeros(n, dtype=RESULT_DTYPE)
        if cells is None:
            cells = np.zeros(n if want_cells else 0, dtype=np.uint32)
        assert out.dtype == RESULT_DTYPE and len(out) == n and out.flags.c_contiguous
        self._check(lib().bsw_extend_batch_flat(self.handle, C.byref(params), qbuf.ctypes.data, qoff.ctypes.data,
                                                tbuf.ctypes.data, toff.ctypes.data, h0.ctypes.data, w.ctypes.data, n,
                                                out.ctypes.data, cells.ctypes.data if want_cells else None))
        return out, cells

    extend_batch_flat = sw_extend_batch

    def sw_extend_tasks(self, params: Params, queries, targets, h0, w):
        """Array-of-task-records layout (bsw_task): one pointer pair per task."""
        n = len(queries)
        qs = [_u8(q) for q in queries]
        ts = [_u8(t) for t in targets]
        tasks = (Task * n)()
        h0 = np.broadcast_to(np.asarray(h0, dtype=np.int32), (n,))
        w = np.broadcast_to(np.asarray(w, dtype=np.int32), (n,))
        for i in range(n):
            tasks[i].query, tasks[i].target = qs[i].ctypes.data, ts[i].ctypes.data
            tasks[i].qlen, tasks[i].tlen, tasks[i].h0, tasks[i].w = len(qs[i]), len(ts[i]), int(h0[i]), int(w[i])
        out = np.zeros(n, dtype=RESULT_DTYPE)
        self._check(lib().bsw_extend_batch(self.handle, C.byref(params), tasks, n, out.ctypes.data))
        return out

    # ---- level 2: proc_element ----
    def proc_element_batch(self, params2: Params2, seeds):
        """seeds: list of dict(q_left,q_right,t_left,t_right,init_score,qbeg,h0,id); left flanks already reversed."""
        n = len(seeds)
        tasks, keep = make_seed_tasks(seeds)
        out = np.zeros(n, dtype=ALN_DTYPE)
        self._check(lib().bsw_chain2aln_batch(self.handle, C.byref(params2), tasks, n, out.ctypes.data))
        del keep
        return out

    chain2aln_batch = proc_element_batch

    # ---- level 3: sw_pe_array over the FPGA wire format ----
    def pe_array_batch(self, tbb_words):
        tbb = np.ascontiguousarray(tbb_words, dtype=np.uint32)
        if tbb.size != TBB_WORDS:
            raise BswError(BSW_EWIRE, "a TBB image is 65536 u32")
        rbb = np.zeros(RBB_WORDS, dtype=np.uint32)
        nres = C.c_int(0)
        self._check(lib().bsw_fpga_batch(self.handle, tbb.ctypes.data, rbb.ctypes.data, C.byref(nres)))
        return rbb, nres.value

    fpga_batch = pe_array_batch

    # ---- ksw_global2: banded global alignment + CIGAR ----
    def global_batch(self, params: Params, queries, targets, w, max_ops: int = 256):
        """queries / targets: lists of uint8 arrays; w: per-task band.  Returns (score int32[n], list of uint32 CIGAR arrays)."""
        n = len(queries)
        qs = [_u8(q) for q in queries]; ts = [_u8(t) for t in targets]
        tasks = (GlobalTask * n)()
        for i in range(n):
            tasks[i].query, tasks[i].target = qs[i].ctypes.data, ts[i].ctypes.data
            tasks[i].qlen, tasks[i].tlen, tasks[i].w = len(qs[i]), len(ts[i]), int(w[i])
        score = np.zeros(n, dtype=np.int32); ncig = np.zeros(n, dtype=np.int32); cig = np.zeros(n * max_ops, dtype=np.uint32)
        self._check(lib().bsw_global_batch(self.handle, C.byref(params), tasks, n, max_ops, score.ctypes.data, ncig.ctypes.data, cig.ctypes.data))
        return score, [cig[i * max_ops: i * max_ops + ncig[i]].copy() for i in range(n)]

    # ---- async pair ----
    def submit(self, params: Params, tasks, n, out):
        t = Ticket()
        self._check(lib().bsw_submit(self.handle, C.byref(params), tasks, n, out.ctypes.data, C.byref(t)))
        return t

    def poll(self, ticket) -> int:
        return lib().bsw_poll(self.handle, C.byref(ticket))

    def wait(self, ticket):
        self._check(lib().bsw_wait(self.handle, C.byref(ticket)))

    # ---- measurement ----
    def resident(self, params: Params, qbuf, qoff, tbuf, toff, h0, w) -> Resident:
        qbuf, qoff, tbuf, toff, h0, w, n = _flat_args(qbuf, qoff, tbuf, toff, h0, w)
        h = C.c_void_p()
        self._check(lib().bsw_resident_create(self.handle, C.byref(params), qbuf.ctypes.data, qoff.ctypes.data,
                                              tbuf.ctypes.data, toff.ctypes.data, h0.ctypes.data, w.ctypes.data, n,
                                              C.byref(h)))
        return Resident(self, h)

    def stats(self) -> dict:
        s = Stats()
        self._check(lib().bsw_get_stats(self.handle, C.byref(s)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def reset_stats(self):
        self._check(lib().bsw_reset_stats(self.handle))

    def measure_int_peak(self, device_index: int = 0) -> dict:
        p = IntPeak()
        self._check(lib().bsw_measure_int_peak(self.handle, device_index, C.byref(p)))
        return {k: getattr(p, k) for k, _ in IntPeak._fields_}


def make_seed_tasks(seeds):
    """Build a (SeedTask * n) array from dicts; returns (array, keepalive list of numpy buffers)."""
    n = len(seeds)
    tasks = (SeedTask * n)()
    keep = []
    for i, s in enumerate(seeds):
        ql, qr, tl, tr = _u8(s["q_left"]), _u8(s["q_right"]), _u8(s["t_left"]), _u8(s["t_right"])
        keep += [ql, qr, tl, tr]
        tasks[i].q_left, tasks[i].q_right = ql.ctypes.data, qr.ctypes.data
        tasks[i].t_left, tasks[i].t_right = tl.ctypes.data, tr.ctypes.data
        tasks[i].qlen[0], tasks[i].qlen[1] = len(ql), len(qr)
        tasks[i].tlen[0], tasks[i].tlen[1] = len(tl), len(tr)
        tasks[i].init_score, tasks[i].qbeg, tasks[i].h0 = int(s["init_score"]), int(s["qbeg"]), int(s["h0"])
        tasks[i].id = int(s.get("id", i))
    return tasks, keep


def tbb_encode(params2: Params2, seeds) -> np.ndarray:
    """Host side of the AFU contract: build the 65536-word task batch buffer image."""
    tasks, keep = make_seed_tasks(seeds)
    tbb = np.zeros(TBB_WORDS, dtype=np.uint32)
    rc = lib().bsw_tbb_encode(C.byref(params2), tasks, len(seeds), tbb.ctypes.data)
    del keep
    if rc != BSW_OK:
        raise BswError(rc, "bsw_tbb_encode")
    return tbb


def fpga_envelope(tbb_words):
    """(tasks outside the FPGA's exact 8-bit envelope, index of the first one or -1) for a TBB image."""
    tbb = np.ascontiguousarray(tbb_words, dtype=np.uint32)
    n, first = C.c_int(0), C.c_int(-1)
    rc = lib().bsw_fpga_envelope(tbb.ctypes.data, C.byref(n), C.byref(first))
    if rc != BSW_OK:
        raise BswError(rc, "bsw_fpga_envelope")
    return n.value, first.value


def rbb_decode(rbb_words, n: int) -> np.ndarray:
    rbb = np.ascontiguousarray(rbb_words, dtype=np.uint32)
    out = np.zeros(n, dtype=ALN_DTYPE)
    rc = lib().bsw_rbb_decode(rbb.ctypes.data, n, out.ctypes.data)
    if rc != BSW_OK:
        raise BswError(rc, "bsw_rbb_decode")
    return out


# ------------------------------------------------------------------------------------------ synthetic workloads
_synth = None


def synth_lib() -> C.CDLL:
    global _synth
    if _synth is None:
        if not os.path.exists(SYNTH_PATH):
            raise RuntimeError(f"{SYNTH_PATH} is missing: run __graft_entry__.build() first")
        S = C.CDLL(SYNTH_PATH)
        S.bsw_synth_shapes.argtypes = [C.POINTER(SynthCfg), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        S.bsw_synth_shapes.restype = None
        S.bsw_synth_fill.argtypes = [C.POINTER(SynthCfg), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        S.bsw_synth_fill.restype = None
        _synth = S
    return _synth


# BASELINE.json configs.  n = number of extension tasks of the full configuration.
WORKLOADS = {
    "cfg1_101bp": dict(n=100_000, read_len=(101, 101), seed=(19, 60), sub=0.01, ins=0.0, dele=0.0, unrelated=0.0, w=100),
    "cfg2_150bp": dict(n=1_000_000, read_len=(150, 150), seed=(19, 60), sub=0.005, ins=0.00025, dele=0.00025, unrelated=0.0, w=100),
    "cfg3_mixed": dict(n=1_000_000, read_len=(50, 250), seed=(19, 60), sub=0.04, ins=0.005, dele=0.005, unrelated=0.10, w=100),
    "cfg4_long": dict(n=20_000, long=True, qlen=(1000, 10000), h0=(19, 200), sub=0.01, ins=0.11, dele=0.03, unrelated=0.0, w=500),
    "cfg5_sweep": dict(n=100_000_000, read_len=(150, 150), seed=(19, 60), sub=0.005, ins=0.00025, dele=0.00025, unrelated=0.0, w=100),
}


def synth_cfg(name: str, seed: int = 1, n_frac: float = 0.0) -> SynthCfg:
    d = WORKLOADS[name]
    c = SynthCfg()
    c.read_len_min, c.read_len_max = d.get("read_len", (0, 0))
    c.seed_min, c.seed_max = d.get("seed", (0, 0))
    c.long_mode = 1 if d.get("long") else 0
    c.qlen_min, c.qlen_max = d.get("qlen", (0, 0))
    c.h0_min, c.h0_max = d.get("h0", (0, 0))
    c.w, c.a, c.o = d["w"], 1, 6
    c.sub, c.ins, c.del_ = d["sub"], d["ins"], d["dele"]
    c.unrelated_frac, c.n_frac, c.seed = d["unrelated"], n_frac, seed
    return c


def synth_tasks(name: str, n: int, first: int = 0, seed: int = 1, n_frac: float = 0.0, qbuf=None, tbuf=None):
    """Tasks [first, first+n) of a named workload: dict(qbuf,qoff,tbuf,toff,h0,w) in the flat level-1 layout.
    qbuf / tbuf: optional caller-owned uint8 buffers to fill (a streaming host reuses its registered batch buffers)."""
    S = synth_lib()
    cfg = synth_cfg(name, seed, n_frac)
    qlen = np.zeros(n, dtype=np.int32)
    tlen = np.zeros(n, dtype=np.int32)
    h0 = np.zeros(n, dtype=np.int32)
    S.bsw_synth_shapes(C.byref(cfg), first, n, qlen.ctypes.data, tlen.ctypes.data, h0.ctypes.data)
    qoff = np.zeros(n + 1, dtype=np.int64)
    toff = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(qlen, out=qoff[1:])
    np.cumsum(tlen, out=toff[1:])
    if qbuf is None or tbuf is None or qbuf.size < int(qoff[-1]) + 8 or tbuf.size < int(toff[-1]) + 8:
        qbuf = np.zeros(int(qoff[-1]) + 8, dtype=np.uint8)
        tbuf = np.zeros(int(toff[-1]) + 8, dtype=np.uint8)
    S.bsw_synth_fill(C.byref(cfg), first, n, qoff.ctypes.data, toff.ctypes.data, qbuf.ctypes.data, tbuf.ctypes.data)
    w = np.full(n, cfg.w, dtype=np.int32)
    return dict(qbuf=qbuf, qoff=qoff, tbuf=tbuf, toff=toff, h0=h0, w=w, n=n)


# ------------------------------------------------------------------------------------------ CPU emulation (tests only)
_emu = None


def emu_lib() -> C.CDLL:
    """TEST ONLY: K1's lane function + scheduler compiled for the host (csrc/emu.cpp).  Never used by Context."""
    global _emu
    if _emu is None:
        if not os.path.exists(EMU_PATH):
            raise RuntimeError(f"{EMU_PATH} is missing: run __graft_entry__.build() first")
        E = C.CDLL(EMU_PATH)
        vp = C.c_void_p
        E.bsw_emu_extend_batch_flat.argtypes = [C.POINTER(Params), C.c_int, vp, vp, vp, vp, vp, vp, C.c_size_t, vp, vp, vp]
        E.bsw_emu_chain2aln.argtypes = [C.POINTER(Params2), C.c_int, vp, C.c_size_t, vp]
        E.bsw_emu_chain2aln_wire.argtypes = [C.POINTER(Params2), C.c_int, vp, C.c_size_t, vp, vp]
        _emu = E
    return _emu


def emu_extend_batch(params: Params, qbuf, qoff, tbuf, toff, h0, w, variant: int = 1):
    qbuf, qoff, tbuf, toff, h0, w, n = _flat_args(qbuf, qoff, tbuf, toff, h0, w)
    out = np.zeros(n, dtype=RESULT_DTYPE)
    cells = np.zeros(n, dtype=np.uint32)
    info = np.zeros(4, dtype=np.int64)
    rc = emu_lib().bsw_emu_extend_batch_flat(C.byref(params), variant, qbuf.ctypes.data, qoff.ctypes.data,
                                             tbuf.ctypes.data, toff.ctypes.data, h0.ctypes.data, w.ctypes.data, n,
                                             out.ctypes.data, cells.ctypes.data, info.ctypes.data)
    if rc != BSW_OK:
        raise BswError(rc, "emulation")
    return out, cells, info


def emu_chain2aln(params2: Params2, seeds, variant: int = 1, wire_gaps=None):
    """TEST ONLY: level 2 through the host scheduler + the K3 lane function compiled for the CPU.
    wire_gaps: (n, 4) max_ins/max_del per side as a TBB carries them = the mode bsw_fpga_batch runs in."""
    tasks, keep = make_seed_tasks(seeds)
    out = np.zeros(len(seeds), dtype=ALN_DTYPE)
    if wire_gaps is not None:
        g = np.ascontiguousarray(wire_gaps, dtype=np.int32).reshape(len(seeds), 4)
        rc = emu_lib().bsw_emu_chain2aln_wire(C.byref(params2), variant, tasks, len(seeds), g.ctypes.data, out.ctypes.data)
    else:
        rc = emu_lib().bsw_emu_chain2aln(C.byref(params2), variant, tasks, len(seeds), out.ctypes.data)
    del keep
    if rc != BSW_OK:
        raise BswError(rc, "emulation (level 2)")
    return out
