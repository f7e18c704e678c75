"""A/B of the e2e batch call (host buffers in, results out) over library options.
   python tools/e2e_ab.py [workload] [n]    env: THREADS="16,4" REG=1 (register the base buffers) CHUNKS="16384,32768"
Prints ms per call (min / median of 9) and the library's own per-call host time for every option combination."""
import os, sys, time, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bsw_b200 as B

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2_150bp"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
threads = [int(x) for x in os.environ.get("THREADS", "16,4").split(",")]
chunks = [int(x) for x in os.environ.get("CHUNKS", "16384").split(",")]
plans = [int(x) for x in os.environ.get("PLANS", "0,1").split(",")]
raws = [int(x) for x in os.environ.get("RAWS", "0").split(",")]
slots = [int(x) for x in os.environ.get("SLOTS", "2").split(",")]
ctx = B.Context(); p = B.make_params()
t = B.synth_tasks(wl, n)
flat = (t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
out = np.zeros(n, dtype=B.RESULT_DTYPE)
ref = None
if os.environ.get("REG", "1") != "0":
    ctx.register_host(t["qbuf"]); ctx.register_host(t["tbuf"])
if os.environ.get("OUTREG", "1") != "0":
    ctx.register_host(out)
for th, ch, dp, raw, sl in itertools.product(threads, chunks, plans, raws, slots):
    ctx.set_option("slots", sl); ctx.set_option("host_threads", th); ctx.set_option("chunk_tasks", ch); ctx.set_option("device_plan", dp); ctx.set_option("raw_inputs", raw)
    for _ in range(3): ctx.sw_extend_batch(p, *flat, want_cells=False, out=out)
    if ref is None: ref = out.copy()
    assert np.array_equal(ref, out), "results changed with the options"
    ctx.reset_stats()
    ts = []
    for _ in range(9):
        t0 = time.perf_counter(); ctx.sw_extend_batch(p, *flat, want_cells=False, out=out); ts.append((time.perf_counter() - t0) * 1e3)
    st = ctx.stats(); ts.sort()
    print("%s n=%d threads %2d chunk %6d device_plan %d raw %d slots %d: min %.2f median %.2f ms  host %.2f ms/call  h2d %.0f MB launches %d"
          % (wl, n, th, ch, dp, raw, sl, ts[0], ts[4], st["pack_ms"] / 9, st["h2d_bytes"] / 9e6, st["kernel_launches"] // 9), flush=True)
