"""Kernel time of one resident plan against the batch size (how well a single chunk fills the GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bsw_b200 as B
ctx = B.Context(); p = B.make_params()
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2_150bp"
for n in (4096, 8192, 16384, 32768, 65536, 131072, 262144, 1000000):
    t = B.synth_tasks(wl, n)
    r = ctx.resident(p, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    for _ in range(3): r.run()
    best = min(r.run()[0] for _ in range(8)); ms, cells, nl = r.run()
    print("n %7d: %.3f ms  %.1f GCUPS  launches %d  (%.1f ns/task)" % (n, best, cells / best * 1e-6, nl, best * 1e6 / n), flush=True)
    r.free() if hasattr(r, "free") else None
