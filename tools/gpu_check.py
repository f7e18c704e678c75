"""Quick on-GPU parity + timing probe (development tool; the judged checks live in tests/ and bench.py)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bsw_b200 as B
import oracle as O


def parity(ctx, name, n, variant=1, n_frac=0.0, seed=1, opts=None, **pk):
    t = B.synth_tasks(name, n, seed=seed, n_frac=n_frac)
    p = B.make_params(**pk); po = O.make_params(**pk)
    ctx.set_option("variant", variant)
    for k, v in (opts or {}).items():
        ctx.set_option(k, v)
    t0 = time.time()
    ro, co = O.extend_batch(po, t['qbuf'], t['qoff'], t['tbuf'], t['toff'], t['h0'], t['w'], variant=variant)
    t1 = time.time()
    rg, cg = ctx.sw_extend_batch(p, t['qbuf'], t['qoff'], t['tbuf'], t['toff'], t['h0'], t['w'])
    t2 = time.time()
    bad = np.nonzero(ro != rg)[0]; badc = np.nonzero(co.astype(np.int64) != cg.astype(np.int64))[0]
    print(f"[parity] {name} n={n} v={variant} nfrac={n_frac} opts={opts} {pk}: mismatch={len(bad)} cells_mismatch={len(badc)} "
          f"oracle={t1-t0:.2f}s gpu_e2e={t2-t1:.3f}s cells={int(co.sum())}", flush=True)
    for i in bad[:3]:
        print("   task", i, "oracle", ro[i], "gpu", rg[i], "qlen", t['qoff'][i+1]-t['qoff'][i], "tlen", t['toff'][i+1]-t['toff'][i], "h0", t['h0'][i])
    ctx.set_option("force_kernel", 0)
    return len(bad) + len(badc)


def timing(ctx, name, n, reps=3, opts=None):
    t = B.synth_tasks(name, n)
    p = B.make_params()
    for k, v in (opts or {}).items():
        ctx.set_option(k, v)
    r = ctx.resident(p, t['qbuf'], t['qoff'], t['tbuf'], t['toff'], t['h0'], t['w'])
    best = None
    for _ in range(reps):
        ms, cells, nl = r.run()
        best = ms if best is None else min(best, ms)
    r.free()
    print(f"[timing] {name} n={n} opts={opts}: kernel {best:.3f} ms, {cells} cells, {nl} launches -> {cells/best*1e-6:.1f} GCUPS, {n/best*1e-3:.2f} Mtasks/s", flush=True)
    t0 = time.time()
    ctx.sw_extend_batch(p, t['qbuf'], t['qoff'], t['tbuf'], t['toff'], t['h0'], t['w'])
    t1 = time.time()
    ctx.reset_stats()
    ctx.sw_extend_batch(p, t['qbuf'], t['qoff'], t['tbuf'], t['toff'], t['h0'], t['w'])
    t2 = time.time()
    print(f"[e2e] first {t1-t0:.3f}s second {t2-t1:.3f}s  stats {ctx.stats()}", flush=True)
    ctx.set_option("force_kernel", 0)


if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "sweep"):
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    ctx = B.Context()
    fails = 0
    if what in ("all", "peak"):
        print("[int peak]", json.dumps(ctx.measure_int_peak()), flush=True)
    if what in ("all", "parity"):
        fails += parity(ctx, "cfg2_150bp", 20000)
        fails += parity(ctx, "cfg3_mixed", 20000)
        fails += parity(ctx, "cfg3_mixed", 20000, n_frac=0.01)
        fails += parity(ctx, "cfg3_mixed", 20000, variant=2)
        fails += parity(ctx, "cfg3_mixed", 5000, opts={"force_kernel": 2})
        fails += parity(ctx, "cfg3_mixed", 5000, n_frac=0.01, opts={"force_kernel": 2})
        fails += parity(ctx, "cfg4_long", 64)
        fails += parity(ctx, "cfg3_mixed", 5000, opts={"force_kernel": 2}, o_del=4, e_del=2, o_ins=7, e_ins=1)
    if what == "long":
        for opts in ({"k2_narrow": 0}, {"k2_narrow": 1, "k2_warps": 1}, {"k2_warps": 4}):
            timing(ctx, "cfg4_long", 20000, reps=2, opts=opts)
        ctx.set_option("k2_narrow", 1); ctx.set_option("k2_warps", 1)
    if what in ("all", "timing"):
        timing(ctx, "cfg2_150bp", 1000000)
        timing(ctx, "cfg3_mixed", 200000)
        timing(ctx, "cfg3_mixed", 50000, opts={"force_kernel": 2})
        timing(ctx, "cfg4_long", 2000)
    print("FAILS", fails)
    sys.exit(1 if fails else 0)


def e2e_sweep():
    import torch
    ctx = B.Context()
    t = B.synth_tasks("cfg2_150bp", 1000000)
    p = B.make_params()
    flat = (t['qbuf'], t['qoff'], t['tbuf'], t['toff'], t['h0'], t['w'])
    x = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); y = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); y.copy_(x, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
        x.copy_(y, non_blocking=True); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"[pcie] H2D {0.268/(t1-t0):.1f} GB/s  D2H {0.268/(t2-t1):.1f} GB/s", flush=True)
    for threads in (16, 8):
        for chunk in (8192, 16384, 32768, 65536, 131072):
            ctx.set_option("chunk_tasks", chunk); ctx.set_option("host_threads", threads)
            for _ in range(3):
                ctx.sw_extend_batch(p, *flat, want_cells=False)
            ctx.reset_stats()
            t0 = time.perf_counter()
            for _ in range(5):
                ctx.sw_extend_batch(p, *flat, want_cells=False)
            dt = (time.perf_counter() - t0) / 5
            st = ctx.stats()
            print(f"[e2e] threads={threads} chunk={chunk}: {dt*1e3:.2f} ms/step  pack_ms/worker={st['pack_ms']/5:.2f} kernel_ms(sum)={st['kernel_ms']/5:.2f} launches={st['kernel_launches']//5}", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "sweep":
    e2e_sweep()
