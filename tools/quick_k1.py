import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bsw_b200 as B
ctx = B.Context(); p = B.make_params()
for name, n in (("cfg2_150bp", 1000000), ("cfg3_mixed", 1000000), ("cfg3_mixed", 200000), ("cfg1_101bp", 100000)):
    t = B.synth_tasks(name, n)
    r = ctx.resident(p, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    best = min(r.run()[0] for _ in range(6)); ms, cells, nl = r.run(); r.free()
    print("%s n=%d: %.3f ms %.1f GCUPS launches=%d" % (name, n, best, cells / best * 1e-6, nl), flush=True)
