"""One resident K2 pass over n cfg4 tasks (for ncu: the kernel of interest is k2_extend_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bsw_b200 as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
ctx = B.Context()
t = B.synth_tasks("cfg4_long", n)
r = ctx.resident(B.make_params(), t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
ms, cells, nl = r.run()
print(f"n={n} {ms:.2f} ms {cells / ms * 1e-6:.1f} GCUPS cells {cells}")
