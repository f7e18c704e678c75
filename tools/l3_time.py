"""Level 3 (bsw_fpga_batch: one TBB image of 800 seed tasks in, one RBB image out) on ONE context: latency of a single image,
its pieces, and throughput with 1 / 2 / 4 / 8 caller threads (the FPGA keeps four PE arrays busy: batch_manager.v:397-562).
The threads call the C ABI directly (ctypes releases the GIL), so what is measured is the library, not Python.
   python tools/l3_time.py > profiles/r02_l3_concurrency.txt"""
import sys, os, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bsw_b200 as B
from helpers import seeds_from_flat
ctx = B.Context()
P2 = B.make_params2(B.make_params(zdrop=0), w=100, pen_clip5=5, pen_clip3=5)
t = B.synth_tasks("cfg1_101bp", 1600, seed=30)
seeds = seeds_from_flat(t, 800, unset_score_every=4)
im = B.tbb_encode(P2, seeds)
for _ in range(5): ctx.pe_array_batch(im)
os.environ["BSW_TRACE"] = "1"
ctx.pe_array_batch(im)
os.environ.pop("BSW_TRACE")
# pieces: level 2 directly
ts=[]
for _ in range(50):
    t0=time.perf_counter(); ctx.proc_element_batch(P2, seeds); ts.append((time.perf_counter()-t0)*1e3)
print("level2 call incl python marshalling: median %.3f ms" % sorted(ts)[25])
tasks, keep = B.make_seed_tasks(seeds)
import ctypes as C
out = np.zeros(800, dtype=B.ALN_DTYPE)
L = B.lib()
ts=[]
for _ in range(50):
    t0=time.perf_counter(); L.bsw_chain2aln_batch(ctx.handle, C.byref(P2), tasks, 800, out.ctypes.data); ts.append((time.perf_counter()-t0)*1e3)
print("bsw_chain2aln_batch alone: median %.3f ms" % sorted(ts)[25])
rbb = np.zeros(4096, np.uint32); n = C.c_int(0)
ts=[]
for _ in range(50):
    t0=time.perf_counter(); L.bsw_fpga_batch(ctx.handle, im.ctypes.data, rbb.ctypes.data, C.byref(n)); ts.append((time.perf_counter()-t0)*1e3)
print("bsw_fpga_batch alone: median %.3f ms" % sorted(ts)[25])
def w():
    r = np.zeros(4096, np.uint32); m = C.c_int(0)
    for _ in range(100): L.bsw_fpga_batch(ctx.handle, im.ctypes.data, r.ctypes.data, C.byref(m))
for nt in (1,2,4,8):
    th=[threading.Thread(target=w) for _ in range(nt)]
    t0=time.perf_counter()
    for x in th: x.start()
    for x in th: x.join()
    dt=time.perf_counter()-t0
    print("%d threads: %.3f ms per image per thread, %.2f M seed tasks/s" % (nt, dt/100*1e3, nt*100*800/dt*1e-6))
