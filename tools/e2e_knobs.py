"""e2e (host buffers in, results out) time of the 1 M task batch against the host-pipeline knobs.
Parent: one child per environment combination; child: sweeps the set_option knobs."""
import sys, os, time, subprocess, itertools
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import bsw_b200 as B
    n = 1000000
    ctx = B.Context(); p = B.make_params()
    t = B.synth_tasks(sys.argv[2], n)
    flat = (t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    out = np.zeros(n, dtype=B.RESULT_DTYPE)
    if os.environ.get("REG"): ctx.register_host(t["qbuf"]); ctx.register_host(t["tbuf"])      # raw mode: no host staging
    tag = " ".join("%s=%s" % (k, os.environ.get(k)) for k in ("CUDA_DEVICE_MAX_CONNECTIONS", "BSW_BUCKET_PCT", "BSW_SIDE_STREAMS"))
    for slots in (2, 3):
        for chunk in (8192, 12288, 16384, 24576):
            ctx.set_option("slots", slots); ctx.set_option("chunk_tasks", chunk)
            for _ in range(3): ctx.sw_extend_batch(p, *flat, want_cells=False, out=out)
            ts = []
            for _ in range(9):
                t0 = time.perf_counter(); ctx.sw_extend_batch(p, *flat, want_cells=False, out=out); ts.append((time.perf_counter() - t0) * 1e3)
            ts.sort()
            print("%s slots %d chunk %6d: min %.2f median %.2f ms" % (tag, slots, chunk, ts[0], ts[len(ts) // 2]), flush=True)
else:
    wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2_150bp"
    for conn, pct, side in itertools.product(("8", "32"), ("130", "200", "100000"), ("1", "3")):
        env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS=conn, BSW_BUCKET_PCT=pct, BSW_SIDE_STREAMS=side)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child", wl], env=env)
