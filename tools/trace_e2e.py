"""Per-worker timeline (BSW_TRACE) of one e2e call on the 1 M x 150 bp batch; knobs come from the environment."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bsw_b200 as B
import numpy as np
ctx = B.Context()
n = int(os.environ.get("N", "1000000"))
t = B.synth_tasks(os.environ.get("WL", "cfg2_150bp"), n)
p = B.make_params()
flat = (t['qbuf'], t['qoff'], t['tbuf'], t['toff'], t['h0'], t['w'])
if os.environ.get("CHUNK"): ctx.set_option("chunk_tasks", int(os.environ["CHUNK"]))
if os.environ.get("SLOTS"): ctx.set_option("slots", int(os.environ["SLOTS"]))
if os.environ.get("GPU_TIMELINE"): ctx.set_option("kernel_timing", 1)     # adds the per-chunk GPU timeline to the trace
out = np.zeros(n, dtype=B.RESULT_DTYPE)
if os.environ.get("REG"): ctx.register_host(t["qbuf"]); ctx.register_host(t["tbuf"])      # raw mode: no host staging
for _ in range(4): ctx.sw_extend_batch(p, *flat, want_cells=False, out=out)
os.environ["BSW_TRACE"] = "1"
t0 = time.perf_counter(); ctx.sw_extend_batch(p, *flat, want_cells=False, out=out); print("total ms", (time.perf_counter() - t0) * 1e3)
