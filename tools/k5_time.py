"""K5 (32-bit rows, one warp per task) against K2 on the same long-read tasks: wall time of the blocking flat call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bsw_b200 as B

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
ctx = B.Context()
p = B.make_params()
for wl, m in (("cfg4_long", n), ("cfg2_150bp", 50 * n)):
    t = B.synth_tasks(wl, m, seed=5)
    for wide in (1, 2):
        ctx.set_option("wide", wide)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            out, cells = ctx.sw_extend_batch(p, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
            best = min(best, time.perf_counter() - t0)
        c = int(cells.astype(np.int64).sum())
        print(f"{wl} n={m} wide={wide} ({'K5 only' if wide == 2 else 'K1/K2'}): {best * 1e3:.1f} ms  {c / best * 1e-9:.1f} GCUPS e2e (cells {c})")
ctx.set_option("wide", 1)
