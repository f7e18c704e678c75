import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bsw_b200 as B
ctx = B.Context(); p = B.make_params()
for k in ("k2_narrow", "k2_warps", "k2_min_qlen"):
    if os.environ.get(k.upper()): ctx.set_option(k, int(os.environ[k.upper()]))
t = B.synth_tasks("cfg4_long", 20000)
flat = (t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
out = np.zeros(20000, dtype=B.RESULT_DTYPE)
for k in range(8):
    ctx.reset_stats(); t0 = time.perf_counter(); ctx.sw_extend_batch(p, *flat, want_cells=False, out=out); dt = time.perf_counter() - t0
    st = ctx.stats()
    print("call %d: %.1f ms  (%.1f GCUPS)  launches %d pack/worker %.1f ms" % (k, dt * 1e3, st["cells_band"] / dt * 1e-9, st["kernel_launches"], st["pack_ms"]), flush=True)
