"""K1 under V1 and V2 (upstream BWA) on the bench workload: resident GCUPS, a sample checked against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bsw_b200 as B
import oracle as O
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2_150bp"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
ctx = B.Context()
t = B.synth_tasks(wl, n)
flat = (t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
p, po = B.make_params(), O.make_params()
for variant in (1, 2):
    ctx.set_option("variant", variant)
    r = ctx.resident(p, *flat)
    ms = min(r.run()[0] for _ in range(3)); _, cells, nl = r.run()
    res, cl = r.fetch(n); r.free()
    m = min(n, 100000)
    ro, co = O.extend_batch(po, t["qbuf"], t["qoff"][:m + 1], t["tbuf"], t["toff"][:m + 1], t["h0"][:m], t["w"][:m], variant=variant)
    ok = bool(np.array_equal(ro, res[:m]) and np.array_equal(co, cl[:m].astype(np.int64)))
    print(f"{wl} n={n} V{variant}: {ms:.3f} ms  {cells / ms * 1e-6:.1f} GCUPS  cells {cells}  first {m} tasks bit-exact: {ok}", flush=True)
ctx.set_option("variant", 1)
