"""K2 on cfg4 long reads under the V2 (upstream BWA) recurrence: resident GCUPS, all tasks checked against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bsw_b200 as B
import oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
ctx = B.Context()
for variant in (1, 2):
    for zd in (100, 400):
        p, po = B.make_params(zdrop=zd), O.make_params(zdrop=zd)
        t = B.synth_tasks("cfg4_long", n)
        flat = (t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
        ro, co = O.extend_batch(po, *flat, variant=variant)
        ctx.set_option("variant", variant)
        r = ctx.resident(p, *flat)
        ms = min(r.run()[0] for _ in range(3)); _, cells, nl = r.run()
        res, cl = r.fetch(n); r.free()
        ok = bool(np.array_equal(ro, res) and np.array_equal(co, cl.astype(np.int64)))
        print(f"cfg4_long n={n} V{variant} zdrop={zd}: {ms:.2f} ms  {cells / ms * 1e-6:.1f} GCUPS  cells {cells}  all tasks bit-exact: {ok}", flush=True)
ctx.set_option("variant", 1)
