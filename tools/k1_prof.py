"""One resident pass over n tasks of a workload (for ncu: kernels k0_gather_kernel, k1_extend_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bsw_b200 as B
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg2_150bp"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
ctx = B.Context()
t = B.synth_tasks(wl, n)
r = ctx.resident(B.make_params(), t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
for _ in range(2): r.run()
ms, cells, nl = r.run()
print(f"{wl} n={n}: {ms:.3f} ms {cells / ms * 1e-6:.1f} GCUPS cells {cells} launches {nl}")
