"""Independent full-matrix model of the V1/V2 extension recurrence -- TEST INFRASTRUCTURE ONLY.

Purpose: a second, differently structured statement of what sw_pe_array_sw_extend.v computes,
used to pin oracle/ksw_extend_ref.c (which uses BWA's rolling eh[] array).  Here every row keeps
its own dictionaries H_i[j], E_i[j] indexed by absolute (i, j); nothing is rolled or reused, so a
stale-value or off-by-one slip in the rolling formulation shows up as a mismatch.

Textbook form (sw_pe_array_sw_extend.v:1797-1799,1809,1862-1866,1770-1771,1780-1781):
    H(i,j)   = max(D(i,j) + S(t_i,q_j), E(i,j), F(i,j))         D = H(i-1,j-1) "as visible"
    E(i+1,j) = max(E(i,j) - e_del, max(0, G(i,j) - o_del - e_del))
    F(i,j+1) = max(F(i,j) - e_ins, max(0, G(i,j) - o_ins - e_ins))
with G = H (V1) or G = M = D+S guarded (V2), on the per-row active window [beg_i, end_i).
Pure Python: small cases only.
"""
from __future__ import annotations


def clamp_w(mat, qlen, w, o_del, e_del, o_ins, e_ins, end_bonus):
    mx = max(0, max(int(x) for x in mat))
    max_ins = int(float(qlen * mx + end_bonus - o_ins) / e_ins + 1.0)
    max_ins = max(max_ins, 1)
    max_del = int(float(qlen * mx + end_bonus - o_del) / e_del + 1.0)
    max_del = max(max_del, 1)
    return min(w, max_ins, max_del)


def extend(mat, query, target, h0, w, o_del=6, e_del=1, o_ins=6, e_ins=1, zdrop=100,
           end_bonus=5, variant=1):
    """Returns dict(score,qle,tle,gtle,gscore,max_off,cells)."""
    qlen, tlen = len(query), len(target)
    w = clamp_w(mat, qlen, w, o_del, e_del, o_ins, e_ins, end_bonus)

    def S(t, q):
        return int(mat[5 * int(t) + int(q)])

    # row -1 ("first row fill"): Hm1[j] = H(-1, j) for j = -1 .. qlen-1
    Hm1 = {-1: h0}
    if qlen >= 1:
        Hm1[0] = max(0, h0 - (o_ins + e_ins))
    for j in range(1, qlen):
        Hm1[j] = max(0, Hm1[j - 1] - e_ins)

    def firstcol(i, beg):
        if variant == 1 or beg == 0:
            return max(0, h0 - (o_del + e_del * (i + 1)))
        return 0

    best, best_i, best_j = h0, -1, -1
    max_ie, gscore, max_off = -1, -1, 0
    beg, end = 0, qlen
    prevH = Hm1            # H(i-1, .) as visible to row i: includes the virtual column beg-1
    prevE = {}             # E(i, .) computed by row i-1; missing -> 0
    # V2 only: upstream BWA's zero-scan narrowing may grow `end` by 2, so the next row can read an
    # eh[] slot that the previous row did not write (a stale slot).  V1 never does (asserted in vis()).
    stale_h = {j: Hm1[j - 1] for j in range(0, qlen + 1)}
    stale_e = {j: 0 for j in range(0, qlen + 1)}
    cells = 0
    for i in range(tlen):
        beg = max(beg, i - w)
        end = min(end, i + w + 1, qlen)
        H = {}
        Enext = {}
        H[beg - 1] = firstcol(i, beg)          # virtual left neighbour (what BWA keeps in h1)
        f = 0
        m, mj = 0, -1
        for j in range(beg, end):
            if variant == 1 or (j - 1) in prevH:
                d = prevH[j - 1]
                e = prevE.get(j, 0)
            else:
                d, e = stale_h[j], stale_e[j]
            if variant == 1:
                M = d + S(target[i], query[j])
            else:
                M = d + S(target[i], query[j]) if d != 0 else 0
            h = max(M, e, f)
            H[j] = h
            if not (m > h):
                mj = j
            m = max(m, h)
            g = h if variant == 1 else M
            Enext[j] = max(e - e_del, max(0, g - (o_del + e_del)))
            f = max(f - e_ins, max(0, g - (o_ins + e_ins)))
        cells += max(0, end - beg)
        last = max(beg, end)                   # value of the loop variable j after the loop
        hlast = H[last - 1] if end > beg else H[beg - 1]
        if last == qlen:
            if not (gscore > hlast):
                max_ie, gscore = i, hlast
        if m == 0:
            break
        if m > best:
            best, best_i, best_j = m, i, mj
            max_off = max(max_off, abs(mj - i))
        elif zdrop > 0:
            di, dj = i - best_i, mj - best_j
            if di > dj:
                if best - m - (di - dj) * e_del > zdrop:
                    break
            else:
                if best - m - (dj - di) * e_ins > zdrop:
                    break
        # "visible" H of this row for the next one: eh[j].h == vis(j-1) for j in [beg, end]
        def vis(jm1, H=H, beg=beg, end=end):
            # eh[] entries outside [beg, end] are never consulted by a correct implementation
            assert beg - 1 <= jm1 <= end - 1, (jm1, beg, end)
            return H[jm1]
        def ecur(j, Enext=Enext):
            return Enext.get(j, 0)
        if variant == 1:
            j = mj
            while j >= beg and vis(j - 1) != 0:
                j -= 1
            nbeg = j + 1
            j = mj + 2
            while j <= end and vis(j - 1) != 0:
                j += 1
            nend = j
        else:
            j = beg
            while j < end and vis(j - 1) == 0 and ecur(j) == 0:
                j += 1
            nbeg = j
            j = end
            while j >= nbeg and vis(j - 1) == 0 and ecur(j) == 0:
                j -= 1
            nend = min(j + 2, qlen)
        for j in range(beg, end + 1):
            stale_h[j] = H[j - 1]
            stale_e[j] = Enext.get(j, 0)
        prevH = H
        prevE = Enext
        beg, end = nbeg, nend
    return dict(score=best, qle=best_j + 1, tle=best_i + 1, gtle=max_ie + 1, gscore=gscore,
                max_off=max_off, cells=cells)
