import sys, os, time; sys.path.insert(0,'/root/repo')
import bsw_b200 as B
ctx = B.Context()
t = B.synth_tasks("cfg2_150bp", 1000000)
p = B.make_params()
flat = (t['qbuf'], t['qoff'], t['tbuf'], t['toff'], t['h0'], t['w'])
import numpy as np
out=np.zeros(1000000, dtype=B.RESULT_DTYPE)
for _ in range(3): ctx.sw_extend_batch(p, *flat, want_cells=False, out=out)
os.environ["BSW_TRACE"]="1"
t0=time.perf_counter(); ctx.sw_extend_batch(p, *flat, want_cells=False, out=out); print("total ms", (time.perf_counter()-t0)*1e3)
