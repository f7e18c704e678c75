"""Level 3 (bsw_fpga_batch: one TBB image in, one RBB image out) latency per image."""
import sys, os, time
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import numpy as np
import bsw_b200 as B
from helpers import seeds_from_flat
ctx = B.Context()
t = B.synth_tasks("cfg1_101bp", 1600, seed=30)
seeds = seeds_from_flat(t, 800, unset_score_every=4)
P2 = B.make_params2(B.make_params(zdrop=0), w=100, pen_clip5=5, pen_clip3=5)
tbb = B.tbb_encode(P2, seeds)
for _ in range(5): ctx.pe_array_batch(tbb)
ts = []
for _ in range(50):
    t0 = time.perf_counter(); rbb, n = ctx.pe_array_batch(tbb); ts.append((time.perf_counter() - t0) * 1e3)
ts.sort()
print("TBB image with %d seed tasks: min %.3f median %.3f ms -> %.2f M seeds/s" % (n, ts[0], ts[25], n / ts[25] * 1e-3))
