"""Lane efficiency of K1 tiles (sum of cells / sum over tiles of 32 x the longest lane) against the size of the sort.
Per-task cells come from the oracle, so this runs on the CPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bsw_b200 as B, oracle as O
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3_mixed"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 400000
t = B.synth_tasks(wl, n)
_, cells = O.extend_batch(O.make_params(), t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
cells = cells.astype(np.int64)
qlen = np.diff(t["qoff"]).astype(np.int64); tlen = np.diff(t["toff"]).astype(np.int64); h0 = t["h0"].astype(np.int64)
rows = tlen  # upper bound of rows; actual rows unknown here
def eff(order_fn, chunk):
    tot = 0; used = 0
    for c0 in range(0, n, chunk):
        idx = np.arange(c0, min(n, c0 + chunk))
        o = idx[order_fn(idx)]
        cc = cells[o]
        pad = (-len(cc)) % 32
        cc = np.concatenate([cc, np.zeros(pad, dtype=np.int64)]).reshape(-1, 32)
        tot += int(cc.max(axis=1).sum()) * 32; used += int(cc.sum())
    return used / tot
keys = {
    "current (qlen, tlen/4, h0/2)": lambda i: np.lexsort((-(h0[i] >> 1), -(tlen[i] >> 2), -qlen[i])),
    "qlen*tlen": lambda i: np.argsort(-(qlen[i] * tlen[i]), kind="stable"),
    "(qlen/8, tlen/8, h0)": lambda i: np.lexsort((-h0[i], -(tlen[i] >> 3), -(qlen[i] >> 3))),
    "(h0 bucket, qlen, tlen)": lambda i: np.lexsort((-(tlen[i] >> 2), -qlen[i], -(h0[i] >> 3))),
    "oracle cells (upper bound)": lambda i: np.argsort(-cells[i], kind="stable"),
}
for name, fn in keys.items():
    print("%-30s" % name, "  ".join("chunk %7d: %.3f" % (c, eff(fn, c)) for c in (8192, 16384, 65536, n)), flush=True)
print("-- predictors (sort by estimated cells, descending)")
preds = {
    "tlen*min(qlen,h0)": lambda i: tlen[i] * np.minimum(qlen[i], h0[i]),
    "min(tlen,qlen+h0)*min(qlen,h0+8)": lambda i: np.minimum(tlen[i], qlen[i] + h0[i]) * np.minimum(qlen[i], h0[i] + 8),
    "qlen*min(qlen,h0+qlen/2)": lambda i: qlen[i] * np.minimum(qlen[i], h0[i] + qlen[i] // 2),
    "qlen*(qlen+2*h0)": lambda i: qlen[i] * (qlen[i] + 2 * h0[i]),
    "min(tlen,qlen+h0)*(qlen+2*h0)": lambda i: np.minimum(tlen[i], qlen[i] + h0[i]) * (qlen[i] + 2 * h0[i]),
}
for name, f in preds.items():
    fn = (lambda f: (lambda i: np.argsort(-f(i), kind="stable")))(f)
    print("%-36s" % name, "  ".join("chunk %7d: %.3f" % (c, eff(fn, c)) for c in (8192, 16384, 65536, n)), "  corr %.3f" % np.corrcoef(f(np.arange(n)), cells)[0, 1], flush=True)
# what do cells look like against (qlen, h0)?
for q0 in (40, 80, 120):
    for hh in (10, 30, 60):
        m = (np.abs(qlen - q0) < 5) & (np.abs(h0 - hh) < 5)
        if m.sum() > 20: print("qlen~%d h0~%d: n=%d tlen %.0f cells mean %.0f p10 %.0f p90 %.0f  cells/qlen^2 %.2f" % (q0, hh, m.sum(), tlen[m].mean(), cells[m].mean(), np.percentile(cells[m], 10), np.percentile(cells[m], 90), cells[m].mean() / q0 ** 2))
print("-- combined keys: coarse qlen bucket (shared memory / occupancy), then estimated cells")
est = tlen * np.minimum(qlen, h0)
def smem_waste(order_fn, chunk):
    # mean over tiles of (max qlen in tile) / (mean qlen in tile): row-buffer bytes allocated vs needed
    a = 0; b = 0
    for c0 in range(0, n, chunk):
        idx = np.arange(c0, min(n, c0 + chunk)); o = idx[order_fn(idx)]
        q = qlen[o]; pad = (-len(q)) % 32
        q = np.concatenate([q, np.zeros(pad, dtype=np.int64)]).reshape(-1, 32)
        a += int(q.max(axis=1).sum()) * 32; b += int(q.sum())
    return a / b
for sh in (3, 4, 5, 6, 30):
    fn = (lambda sh: (lambda i: np.lexsort((-est[i], -(qlen[i] >> sh)))))(sh)
    print("(qlen>>%d, est cells)" % sh, "  ".join("chunk %7d: eff %.3f qmax/qmean %.2f" % (c, eff(fn, c), smem_waste(fn, c)) for c in (16384, 65536, n)), flush=True)
fn = keys["current (qlen, tlen/4, h0/2)"]
print("current               ", "  ".join("chunk %7d: eff %.3f qmax/qmean %.2f" % (c, eff(fn, c), smem_waste(fn, c)) for c in (16384, 65536, n)), flush=True)
print("-- lockstep model: lanes advance row by row together; tile time = sum over rows of the widest live lane")
width = np.maximum(cells // np.maximum(tlen, 1), 1)
def eff2(order_fn, chunk):
    tot = 0; used = 0
    for c0 in range(0, n, chunk):
        idx = np.arange(c0, min(n, c0 + chunk)); o = idx[order_fn(idx)]
        pad = (-len(o)) % 32
        R = np.concatenate([tlen[o], np.zeros(pad, dtype=np.int64)]).reshape(-1, 32)
        W = np.concatenate([width[o], np.zeros(pad, dtype=np.int64)]).reshape(-1, 32)
        # sort lanes of every tile by rows descending; running max of width over lanes with at least that many rows
        k = np.argsort(-R, axis=1, kind="stable")
        Rs = np.take_along_axis(R, k, axis=1); Ws = np.take_along_axis(W, k, axis=1)
        Wmax = np.maximum.accumulate(Ws, axis=1)
        nxt = np.concatenate([Rs[:, 1:], np.zeros((Rs.shape[0], 1), dtype=np.int64)], axis=1)
        tot += int(((Rs - nxt) * Wmax).sum()) * 32
        used += int(cells[o].sum())
    return used / tot
wp = np.minimum(qlen, h0)
keys2 = {
    "current (qlen, tlen/4, h0/2)": keys["current (qlen, tlen/4, h0/2)"],
    "(qlen>>4, est)": lambda i: np.lexsort((-est[i], -(qlen[i] >> 4))),
    "(qlen>>4, wp>>2, tlen)": lambda i: np.lexsort((-tlen[i], -(wp[i] >> 2), -(qlen[i] >> 4))),
    "(qlen>>4, wp>>3, tlen)": lambda i: np.lexsort((-tlen[i], -(wp[i] >> 3), -(qlen[i] >> 4))),
    "(qlen>>5, wp>>3, tlen)": lambda i: np.lexsort((-tlen[i], -(wp[i] >> 3), -(qlen[i] >> 5))),
    "(qlen>>4, tlen>>4, wp)": lambda i: np.lexsort((-wp[i], -(tlen[i] >> 4), -(qlen[i] >> 4))),
    "(qlen>>5, tlen>>4, wp)": lambda i: np.lexsort((-wp[i], -(tlen[i] >> 4), -(qlen[i] >> 5))),
    "(qlen>>5, tlen>>5, wp)": lambda i: np.lexsort((-wp[i], -(tlen[i] >> 5), -(qlen[i] >> 5))),
    "(wp>>3, tlen)": lambda i: np.lexsort((-tlen[i], -(wp[i] >> 3))),
}
for name, fn in keys2.items():
    print("%-30s" % name, "  ".join("chunk %7d: cells-eff %.3f lockstep-eff %.3f" % (c, eff(fn, c), eff2(fn, c)) for c in (16384, 65536, n)), flush=True)
