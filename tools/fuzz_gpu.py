"""Randomised parity campaign on the GPU: medium-length adversarial tasks (rows long enough for the 16-cell path, ties,
indels, N, tiny h0 / w), random scoring, V1 and V2, level 1 through the default kernel choice, K2 and K5, level 2 fused,
K4 global alignment.
Usage: [RAW=1] python tools/fuzz_gpu.py [iterations] [first_seed]   (RAW=1: every other iteration uses registered buffers)"""
import sys, os, time
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import numpy as np
import bsw_b200 as B, oracle as O
from helpers import random_small_tasks, oracle_chain2aln

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = B.Context()
fails = 0
ntask_checked = 0
nseed_checked = 0
nglobal_checked = 0
t_start = time.time()
for it in range(iters):
    rng = np.random.default_rng(777000 + seed0 + it)
    qmax, tmax = int(rng.choice([40, 120, 300, 600])), int(rng.choice([60, 200, 400, 800]))
    n = 3000 if qmax <= 300 else 800
    t = random_small_tasks(rng, n, qmax=qmax, tmax=tmax)
    if it % 3 == 0:                                      # sprinkle N
        m = rng.random(len(t["qbuf"])) < 0.01; t["qbuf"] = np.where(m, 4, t["qbuf"]).astype(np.uint8)
        m = rng.random(len(t["tbuf"])) < 0.01; t["tbuf"] = np.where(m, 4, t["tbuf"]).astype(np.uint8)
    if it % 2 == 0:                                      # larger score budgets: wide live windows
        t["h0"] = rng.integers(1, 200, n).astype(np.int32)
    pk = dict() if it % 4 == 0 else dict(o_del=int(rng.integers(0, 8)), e_del=int(rng.integers(1, 4)), o_ins=int(rng.integers(0, 8)),
                                         e_ins=int(rng.integers(1, 4)), zdrop=int(rng.choice([0, 5, 30, 100])), a=int(rng.integers(1, 4)),
                                         b=int(rng.integers(1, 7)), end_bonus=int(rng.integers(0, 10)))
    p, po = B.make_params(**pk), O.make_params(**pk)
    flat = (t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    raw = bool(os.environ.get("RAW")) and it % 2 == 1          # every other iteration through the raw input mode
    if raw:
        ctx.register_host(t["qbuf"]); ctx.register_host(t["tbuf"]); ctx.set_option("raw_inputs", 1)
    for variant in (1, 2):
        ro, co = O.extend_batch(po, *flat, variant=variant)
        for opts in (dict(), dict(force_kernel=2), dict(wide=2)):          # default kernel choice, K2, K5 (32-bit rows)
            ctx.set_option("variant", variant)
            for k, v in opts.items(): ctx.set_option(k, v)
            rg, cg = ctx.sw_extend_batch(p, *flat)
            ctx.set_option("force_kernel", 0); ctx.set_option("variant", 1); ctx.set_option("wide", 1)
            ntask_checked += n
            if not (np.array_equal(ro, rg) and np.array_equal(co.astype(np.int64), cg.astype(np.int64))):
                bad = np.nonzero((ro != rg) | (co.astype(np.int64) != cg.astype(np.int64)))[0]
                print("MISMATCH it=%d variant=%d opts=%s pk=%s first task %d: oracle %s gpu %s" % (it, variant, opts, pk, bad[0], ro[bad[0]], rg[bad[0]]), flush=True)
                fails += 1
    if raw:
        ctx.unregister_host(t["qbuf"]); ctx.unregister_host(t["tbuf"]); ctx.set_option("raw_inputs", 2)
    # level 2 on flank pairs
    seeds = []
    sl = lambda a, off, i: a[off[i]:off[i + 1]]
    for r in range(n // 2):
        l, g = 2 * r, 2 * r + 1
        ql, tl, qr, tr = sl(t["qbuf"], t["qoff"], l), sl(t["tbuf"], t["toff"], l), sl(t["qbuf"], t["qoff"], g), sl(t["tbuf"], t["toff"], g)
        if r % 5 == 1: ql, tl = ql[:0], tl[:0]
        if r % 5 == 2: qr, tr = qr[:0], tr[:0]
        h0 = int(t["h0"][l])
        seeds.append(dict(q_left=ql, q_right=qr, t_left=tl, t_right=tr, init_score=(h0 if len(ql) == 0 else -1), qbeg=len(ql), h0=h0, id=r))
    P2 = B.make_params2(p, w=int(rng.choice([3, 20, 100])), pen_clip5=int(rng.integers(0, 8)), pen_clip3=int(rng.integers(0, 8)))
    want, _ = oracle_chain2aln(O, B, P2, seeds)
    got = ctx.proc_element_batch(P2, seeds)
    if not np.array_equal(want, got):
        bad = np.nonzero(want != got)[0]
        print("MISMATCH L2 it=%d pk=%s first seed %d: oracle %s gpu %s" % (it, pk, bad[0], want[bad[0]], got[bad[0]]), flush=True); fails += 1
    nseed_checked += len(seeds)
    # K4: banded global alignment + CIGAR on the first 300 (query, target) pairs, band = length difference + a random margin
    m = min(300, n)
    qs = [sl(t["qbuf"], t["qoff"], i) for i in range(m)]; ts = [sl(t["tbuf"], t["toff"], i) for i in range(m)]
    ws = [abs(len(a) - len(b)) + int(rng.integers(0, 40)) for a, b in zip(qs, ts)]
    sc, cg = ctx.global_batch(p, qs, ts, ws, max_ops=2048)
    for i in range(m):
        s0, c0 = O.global_align(po, qs[i], ts[i], ws[i], max_cigar=2048)
        if s0 != sc[i] or not np.array_equal(c0, cg[i]):
            print("MISMATCH K4 it=%d pk=%s task %d: oracle %d gpu %d" % (it, pk, i, s0, sc[i]), flush=True); fails += 1; break
    nglobal_checked += m
print("fuzz: %d iterations, %d failures, %.1f s; %d extension results, %d seed records, %d global alignments compared with the oracle"
      % (iters, fails, time.time() - t_start, ntask_checked, nseed_checked, nglobal_checked), flush=True)
sys.exit(1 if fails else 0)
