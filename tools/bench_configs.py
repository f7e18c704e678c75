"""All five BASELINE.json configs on one GPU: resident GCUPS, end-to-end GCUPS, bit-exactness on an oracle sample.
Informational companion of bench.py (which measures configs[1]); writes one JSON line per config."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bsw_b200 as B
import oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import oracle_chain2aln, seeds_from_flat


def run(ctx, label, workload, n, sample, reps=3, **pk):
    t = B.synth_tasks(workload, n)
    p, po = B.make_params(**pk), O.make_params(**pk)
    flat = (t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    r = ctx.resident(p, *flat)
    ms = min(r.run()[0] for _ in range(reps)); _, cells, nl = r.run()
    res, cl = r.fetch(n); r.free()
    out = np.zeros(n, dtype=B.RESULT_DTYPE)
    for _ in range(3):                      # warm-up: the staging buffers of every worker slot grow to this workload's chunks
        ctx.sw_extend_batch(p, *flat, want_cells=False, out=out)
    ts = []
    for _ in range(max(reps, 5)):
        t0 = time.perf_counter()
        ctx.sw_extend_batch(p, *flat, want_cells=False, out=out)
        ts.append(time.perf_counter() - t0)
    e2e = sorted(ts)[len(ts) // 2]          # median
    # oracle on a sample (first `sample` tasks), multi-threaded; also the CPU rate
    s = min(sample, n)
    t0 = time.perf_counter()
    ro, co = O.extend_batch(po, t["qbuf"], t["qoff"][:s + 1], t["tbuf"], t["toff"][:s + 1], t["h0"][:s], t["w"][:s])
    cpu_s = time.perf_counter() - t0
    exact = bool(np.array_equal(ro, res[:s]) and np.array_equal(ro, out[:s]) and np.array_equal(co, cl[:s].astype(np.int64)))
    line = dict(config=label, workload=workload, tasks=n, cells=int(cells), launches=nl, kernel_ms=ms, gcups_resident=cells / ms * 1e-6,
                mtasks_per_s_resident=n / ms * 1e-3, e2e_ms=e2e * 1e3, gcups_e2e=cells / e2e * 1e-9, bit_exact_vs_oracle=exact,
                oracle_sample=s, cpu_gcups_all_cores=float(co.sum()) / cpu_s * 1e-9, cpu_threads=O.max_threads(), scoring=pk or "defaults")
    print(json.dumps(line), flush=True)
    return line


if __name__ == "__main__":
    ctx = B.Context()
    out = []
    only_l2 = len(sys.argv) > 1 and sys.argv[1] == "l2"
    if not only_l2:
      out.append(run(ctx, "cfg1 100k x 101 bp", "cfg1_101bp", 100_000, 100_000))
      out.append(run(ctx, "cfg2 1M x 150 bp", "cfg2_150bp", 1_000_000, 100_000))
      out.append(run(ctx, "cfg3 1M mixed 50-250 bp, 10% unrelated", "cfg3_mixed", 1_000_000, 100_000))
      out.append(run(ctx, "cfg4 20k x 1-10 kb w=500 zdrop=100", "cfg4_long", 20_000, 500, reps=2))
      out.append(run(ctx, "cfg4 20k x 1-10 kb w=500 zdrop=400", "cfg4_long", 20_000, 500, reps=2, zdrop=400))
    # cfg5: the 100M-task sweep is cfg2 streamed in 1M chunks regenerated from (seed, first); here 5 chunks, parity on 1%
    p, po = B.make_params(), O.make_params()
    tot_cells, tot_s, ok = 0, 0.0, True
    outbuf = np.zeros(1_000_000, dtype=B.RESULT_DTYPE)
    for c in range(0 if only_l2 else 5):
        t = B.synth_tasks("cfg5_sweep", 1_000_000, first=c * 1_000_000)
        flat = (t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
        ctx.reset_stats()
        t0 = time.perf_counter(); ctx.sw_extend_batch(p, *flat, want_cells=False, out=outbuf); tot_s += time.perf_counter() - t0
        tot_cells += ctx.stats()["cells_band"]
        ro, _ = O.extend_batch(po, t["qbuf"], t["qoff"][:10001], t["tbuf"], t["toff"][:10001], t["h0"][:10000], t["w"][:10000])
        ok = ok and bool(np.array_equal(ro, outbuf[:10000]))
    if not only_l2:
        line = dict(config="cfg5 sweep (5 of 100 chunks of 1M, streamed through the C ABI)", tasks=5_000_000, cells=int(tot_cells),
                    e2e_ms_per_chunk=tot_s / 5 * 1e3, gcups_e2e=tot_cells / tot_s * 1e-9, mtasks_per_s_e2e=5.0 / tot_s, bit_exact_1pct_sample=ok)
        print(json.dumps(line), flush=True)
    # level 2 (seed tasks): fused K3 vs host-orchestrated
    t = B.synth_tasks("cfg2_150bp", 400_000)
    seeds = seeds_from_flat(t, 200_000, unset_score_every=3)
    P2 = B.make_params2()
    tasks, keep = B.make_seed_tasks(seeds)
    rec = np.zeros(len(seeds), dtype=B.ALN_DTYPE)
    for fused in (1, 0):
        ctx.set_option("fused_l2", fused)
        for _ in range(3):
            B.lib().bsw_chain2aln_batch(ctx.handle, __import__("ctypes").byref(P2), tasks, len(seeds), rec.ctypes.data)
        dts = []
        for _ in range(4):
            t0 = time.perf_counter()
            rc = B.lib().bsw_chain2aln_batch(ctx.handle, __import__("ctypes").byref(P2), tasks, len(seeds), rec.ctypes.data)
            dts.append(time.perf_counter() - t0)
        dt = min(dts)
        want, _ = oracle_chain2aln(O, B, P2, seeds[:20000])
        print(json.dumps(dict(config="level 2 (proc_element) 200k seed tasks from 150 bp reads", fused_l2=fused, rc=rc, e2e_ms=dt * 1e3, all_ms=[round(x * 1e3, 1) for x in dts],
                              mseeds_per_s=len(seeds) / dt * 1e-6, bit_exact_vs_oracle=bool(np.array_equal(want, rec[:20000])))), flush=True)
    ctx.set_option("fused_l2", 1)
