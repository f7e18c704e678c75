cat > /tmp/replay.py <<'PY'
import sys, os, time; sys.path.insert(0, "/root/repo")
import bsw_b200 as B, numpy as np
ctx = B.Context(); n = 1000000
t = B.synth_tasks("cfg2_150bp", n); p = B.make_params()
flat = (t['qbuf'], t['qoff'], t['tbuf'], t['toff'], t['h0'], t['w'])
out = np.zeros(n, dtype=B.RESULT_DTYPE)
tag = " ".join("%s=%s" % (k, os.environ.get(k)) for k in ("CUDA_DEVICE_MAX_CONNECTIONS", "BSW_BUCKET_PCT", "BSW_SIDE_STREAMS"))
for slots in (2, 4):
  for chunk in (16384, 32768):
    ctx.set_option("chunk_tasks", chunk); ctx.set_option("slots", slots)
    os.environ.pop("BSW_REPLAY", None)
    for _ in range(4): ctx.sw_extend_batch(p, *flat, want_cells=False, out=out)
    ts = []
    for _ in range(7):
        t0 = time.perf_counter(); ctx.sw_extend_batch(p, *flat, want_cells=False, out=out); ts.append((time.perf_counter() - t0) * 1e3)
    full = sorted(ts)[3]
    os.environ["BSW_REPLAY"] = "1"
    ts = []
    for _ in range(7):
        t0 = time.perf_counter(); ctx.sw_extend_batch(p, *flat, want_cells=False, out=out); ts.append((time.perf_counter() - t0) * 1e3)
    print("%s slots %d chunk %d: full %.2f ms, GPU-side only (replay) %.2f ms" % (tag, slots, chunk, full, sorted(ts)[3]), flush=True)
PY
for conn in 32 128; do for pct in 130 200 100000; do for side in 0 1 3; do
CUDA_DEVICE_MAX_CONNECTIONS=$conn BSW_BUCKET_PCT=$pct BSW_SIDE_STREAMS=$side timeout 300 python /tmp/replay.py 2>&1 | tail -4
done; done; done
