"""Experiment: K1P upper bound -- every task duplicated so the two tasks of a lane have identical windows."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bsw_b200 as B
n = 500000
t = B.synth_tasks("cfg2_150bp", n)
ql = np.diff(t["qoff"]); tl = np.diff(t["toff"])
idx = np.repeat(np.arange(n), 2)
qoff = np.concatenate([[0], np.cumsum(ql[idx])]).astype(np.int64); toff = np.concatenate([[0], np.cumsum(tl[idx])]).astype(np.int64)
def gather(buf, off, lens):
    out = np.zeros(int(lens[idx].sum()) + 8, np.uint8); pos = 0
    starts = off[:-1][idx]; ln = lens[idx]
    # vectorised gather
    total = int(ln.sum()); rep = np.repeat(starts - np.concatenate([[0], np.cumsum(ln)[:-1]]), ln)
    out[:total] = buf[np.arange(total) + rep]
    return out
qbuf = gather(t["qbuf"], t["qoff"], ql); tbuf = gather(t["tbuf"], t["toff"], tl)
h0 = t["h0"][idx]; w = t["w"][idx]
ctx = B.Context()
p = B.make_params()
for pair in (1, 0):
    ctx.set_option("k1_pair", pair)
    r = ctx.resident(p, qbuf, qoff, tbuf, toff, h0, w)
    best = min(r.run()[0] for _ in range(4)); ms, cells, nl = r.run(); r.free()
    print(f"duplicated tasks pair={pair}: {best:.3f} ms {cells/best*1e-6:.1f} GCUPS launches={nl}", flush=True)
    r = ctx.resident(p, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    best = min(r.run()[0] for _ in range(4)); ms, cells, nl = r.run(); r.free()
    print(f"plain 500k tasks   pair={pair}: {best:.3f} ms {cells/best*1e-6:.1f} GCUPS launches={nl}", flush=True)
