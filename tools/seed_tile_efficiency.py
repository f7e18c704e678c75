"""Lane efficiency of K3 tiles (level 2: one seed per lane, left then right extension) for different sort keys.
Model: tile time = longest left flank + longest right flank (the two phases are separate code, lanes reconverge)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bsw_b200 as B, oracle as O
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
t = B.synth_tasks("cfg2_150bp", 2 * ns)
_, cells = O.extend_batch(O.make_params(), t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
cells = cells.astype(np.int64); qlen = np.diff(t["qoff"]).astype(np.int64); tlen = np.diff(t["toff"]).astype(np.int64); h0 = t["h0"].astype(np.int64)
r = np.arange(ns); kind = r % 4
cl, cr = cells[2 * r].copy(), cells[2 * r + 1].copy(); ql, qr = qlen[2 * r].copy(), qlen[2 * r + 1].copy()
tl, tr = tlen[2 * r].copy(), tlen[2 * r + 1].copy(); hh = h0[2 * r]
cl[kind == 1] = 0; ql[kind == 1] = 0; tl[kind == 1] = 0; cr[kind == 2] = 0; qr[kind == 2] = 0; tr[kind == 2] = 0
el = tl * np.minimum(ql, hh); er = tr * np.minimum(qr, hh + ql)      # right flank starts from the left score ~ h0 + ql
def eff(order_fn, chunk):
    tot = used = 0
    for c0 in range(0, ns, chunk):
        idx = np.arange(c0, min(ns, c0 + chunk)); o = idx[order_fn(idx)]
        pad = (-len(o)) % 32
        L = np.concatenate([cl[o], np.zeros(pad, dtype=np.int64)]).reshape(-1, 32); R = np.concatenate([cr[o], np.zeros(pad, dtype=np.int64)]).reshape(-1, 32)
        tot += int((L.max(axis=1) + R.max(axis=1)).sum()) * 32; used += int(L.sum() + R.sum())
    return used / tot
qm = np.maximum(ql, qr)
keys = {
    "current (max q, (ql+qr)/2, h0/2)": lambda i: np.lexsort((-(hh[i] >> 1), -((ql[i] + qr[i]) >> 1), -qm[i])),
    "(max q>>4, est L+R)": lambda i: np.lexsort((-(el[i] + er[i]), -(qm[i] >> 4))),
    "(max q>>4, est L>>9, est R)": lambda i: np.lexsort((-er[i], -(el[i] >> 9), -(qm[i] >> 4))),
    "(max q>>4, est L>>10, est R)": lambda i: np.lexsort((-er[i], -(el[i] >> 10), -(qm[i] >> 4))),
    "(max q>>5, est L>>10, est R)": lambda i: np.lexsort((-er[i], -(el[i] >> 10), -(qm[i] >> 5))),
    "(est L>>10, est R)": lambda i: np.lexsort((-er[i], -(el[i] >> 10))),
    "oracle (cl>>9, cr)": lambda i: np.lexsort((-cr[i], -(cl[i] >> 9))),
}
for name, fn in keys.items():
    print("%-36s" % name, "  ".join("chunk %6d: %.3f" % (c, eff(fn, c)) for c in (4096, 8192, 32768, ns)), flush=True)
