"""Level 2 (bsw_chain2aln_batch, fused K3) timing on seed tasks cut from 150 bp reads; BSW_TRACE=1 shows the worker timeline."""
import sys, os, time, ctypes
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import numpy as np
import bsw_b200 as B
from helpers import seeds_from_flat
nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
ctx = B.Context()
t = B.synth_tasks("cfg2_150bp", 2 * nseeds)
seeds = seeds_from_flat(t, nseeds, unset_score_every=3)
P2 = B.make_params2()
tasks, keep = B.make_seed_tasks(seeds)
rec = np.zeros(len(seeds), dtype=B.ALN_DTYPE)
call = lambda: B.lib().bsw_chain2aln_batch(ctx.handle, ctypes.byref(P2), tasks, len(seeds), rec.ctypes.data)
for _ in range(4): call()
dts = []
for _ in range(9):
    t0 = time.perf_counter(); rc = call(); dts.append((time.perf_counter() - t0) * 1e3)
dts.sort()
print("%d seeds rc=%d min %.2f median %.2f ms -> %.1f M seeds/s" % (nseeds, rc, dts[0], dts[4], nseeds / dts[4] * 1e-3), flush=True)
