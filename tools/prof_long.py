import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bsw_b200 as B
ctx = B.Context()
t = B.synth_tasks("cfg4_long", 4000)
p = B.make_params()
for sub in (1, 0):
    ctx.set_option("k2_sub", sub)
    r = ctx.resident(p, t['qbuf'], t['qoff'], t['tbuf'], t['toff'], t['h0'], t['w'])
    print(sub, r.run()); r.free()
