"""Small run over every kernel path (K0, planner, K1 fast/matrix/V2, K1R, K2 V1/V2, K3) for compute-sanitizer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import bsw_b200 as B
import oracle as O
from helpers import oracle_chain2aln, seeds_from_flat

ctx = B.Context()
p, po = B.make_params(), O.make_params()
bad = 0


def check(name, n, variant=1, n_frac=0.0, **opts):
    global bad
    t = B.synth_tasks(name, n, seed=3, n_frac=n_frac)
    ctx.set_option("variant", variant)
    for k, v in opts.items():
        ctx.set_option(k, v)
    ro, co = O.extend_batch(po, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"], variant=variant)
    rg, cg = ctx.sw_extend_batch(p, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    ok = np.array_equal(ro, rg) and np.array_equal(co, cg.astype(np.int64))
    print(name, n, variant, n_frac, opts, "OK" if ok else "MISMATCH", flush=True)
    bad += 0 if ok else 1
    for k in opts:
        ctx.set_option(k, {"force_kernel": 0, "k2_narrow": 1, "k2_warps": 1, "k2_min_qlen": 384}[k])
    ctx.set_option("variant", 1)


check("cfg3_mixed", 1500)
check("cfg3_mixed", 1500, n_frac=0.02)
check("cfg3_mixed", 1500, variant=2)
check("cfg3_mixed", 1500, k2_min_qlen=64)
check("cfg3_mixed", 600, force_kernel=2)
check("cfg3_mixed", 600, force_kernel=2, k2_warps=4)
check("cfg3_mixed", 600, variant=2, force_kernel=2)
check("cfg4_long", 6)
t = B.synth_tasks("cfg3_mixed", 1000, seed=4)
seeds = seeds_from_flat(t, 500, unset_score_every=3)
P2 = B.make_params2(w=10)
want, _ = oracle_chain2aln(O, B, P2, seeds)
got = ctx.proc_element_batch(P2, seeds)
print("level 2 fused", "OK" if np.array_equal(want, got) else "MISMATCH", flush=True)
bad += 0 if np.array_equal(want, got) else 1
ctx.close()
print("BAD", bad)
sys.exit(1 if bad else 0)
