"""e2e time per 1 M tasks against the size of one batch call (per-call start/drain cost vs streaming rate)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bsw_b200 as B
ctx = B.Context(); p = B.make_params()
for n in (100_000, 250_000, 1_000_000, 2_000_000, 4_000_000):
    t = B.synth_tasks("cfg2_150bp", n)
    flat = (t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    out = np.zeros(n, dtype=B.RESULT_DTYPE)
    for _ in range(3): ctx.sw_extend_batch(p, *flat, want_cells=False, out=out)
    ts = []
    ctx.reset_stats()
    for _ in range(7):
        t0 = time.perf_counter(); ctx.sw_extend_batch(p, *flat, want_cells=False, out=out); ts.append((time.perf_counter() - t0) * 1e3)
    cells = ctx.stats()["cells_band"] / 7
    ts.sort()
    print("n %8d: median %.2f ms = %.2f ms per 1 M tasks, %.0f GCUPS" % (n, ts[3], ts[3] * 1e6 / n, cells / ts[3] * 1e-6), flush=True)
