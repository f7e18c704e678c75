"""CPU model of K2's packed round (bwa-mem-sw_b200/csrc/bsw_k2.cu: k2_round_packed and the row loop around it), the
intra-task kernel for long reads.  Lane-parallel numpy, the same packed 16x2 words and the same steps:

  * row word {H hi16, E lo16}; every step of the cell one per-half add-max with 16-bit wrap (the DPX instruction);
  * lane l owns CPL = 8 columns of a round (rounds start at j0 & ~7 and repeat), or 12 when the window is 256..383
    columns wide (one round);
  * the F chain in two passes: zero-carry runs per lane, a max-plus prefix scan across the lanes, the carry folded in;
  * dead columns left of j0 zeroed in the row buffer (match bits dropped) instead of a per-cell chain reset;
  * the arg-max over ALL columns of a lane, repaired from the row buffer when a column right of the window wins;
  * V1 narrowing from the lanes' zero bits, V2 (upstream BWA): zero guard as one unsigned minimum, gap opens from M,
    exact boundary stores, first/last non-zero word scan, first-row values restored in ring mode;
  * the 2 048-column ring of long queries (12-column lanes may straddle its end).

Compared with the scalar oracle on every field and the cell count -- so the ALGORITHM is covered by the CPU suite; the
kernel itself is compared with the oracle on the GPU (tests/test_gpu_parity.py).  Every V1 task runs twice: all rows on
the packed round (option k2_narrow = 0), and rows below 64 columns on the registers-only path (k2_narrow_row: one or two
columns per lane from j0, scalar cell, live masks) as the kernel does by default."""
import ctypes as C

import numpy as np
import pytest

from helpers import flat_from_lists, random_small_tasks

U32 = np.uint32
L32 = np.arange(32, dtype=np.int64)
RING = 2048


def s16(x):
    return ((np.asarray(x, dtype=np.int64) & 0xffff) ^ 0x8000) - 0x8000


def pack(hi, lo):
    return (((np.asarray(hi, dtype=np.int64) & 0xffff) << 16) | (np.asarray(lo, dtype=np.int64) & 0xffff)).astype(np.int64)


def hi16(x): return s16(np.asarray(x, dtype=np.int64) >> 16)
def lo16(x): return s16(x)


def add_max2(a, b, c):         # per half: max(wrap16(a + b), c)
    return pack(np.maximum(s16(hi16(a) + hi16(b)), hi16(c)), np.maximum(s16(lo16(a) + lo16(b)), lo16(c)))


def max2(a, b):
    return pack(np.maximum(hi16(a), hi16(b)), np.maximum(lo16(a), lo16(b)))


def k2_model(p, query, target, h0, w, variant, narrow=False):
    mat25 = np.frombuffer(bytes(p.mat), dtype=np.int8).astype(np.int64)
    a, b = int(mat25[0]), -int(mat25[1])
    generic = bool((np.asarray(query) == 4).any() or (np.asarray(target) == 4).any()) or any(
        mat25[5 * i + j] != (a if i == j else -b) for i in range(4) for j in range(4))
    o_del, e_del, o_ins, e_ins, zdrop = p.o_del, p.e_del, p.o_ins, p.e_ins, p.zdrop
    oe_del, oe_ins = o_del + e_del, o_ins + e_ins
    qlen, tlen = len(query), len(target)
    q = np.concatenate([np.asarray(query, dtype=np.int64), np.zeros(64, dtype=np.int64)])      # padding nibbles are 0
    qcap = (qlen + 1 + 255) & ~255
    ring = qcap > RING and 2 * w + 1 + 512 <= RING
    rcap = RING if ring else qcap
    rm = RING - 1 if ring else (1 << 40) - 1
    c_mis = pack(0 if generic else -b, -32768); c_noe_del = pack(-oe_del, 0); c_noe_ins = pack(-oe_ins, 0)
    c_ne_ins = pack(-e_ins, 0); c_eh = pack(-32768, -e_del)
    eh = np.zeros(rcap + 16, dtype=np.int64)
    for j in range(rcap + 16):
        hv = h0 if j == 0 else max(h0 - o_ins - j * e_ins, 0)
        eh[j] = pack(0 if j > qlen else hv, 0)
    zb = np.zeros(rcap + 64, dtype=np.int64)           # one zero / non-zero bit per column (the kernel keeps a byte per 8)
    mx, max_i, max_j, max_ie, gscore, max_off = h0, -1, -1, -1, -1, 0
    beg, end, cells, hiw = 0, qlen, 0, rcap - 1

    def round_packed(rbase, cpl, j0, lim, fc, tb, carry, hcarry, key, single):
        K = np.arange(cpl, dtype=np.int64)
        jl = rbase + cpl * L32
        lo, hi = j0 - jl, lim - jl
        if rbase < j0:
            eh[(rbase + np.arange(j0 - rbase)) & rm] = 0
        col = jl[:, None] + K[None, :]
        wd = np.where((hi >= 0)[:, None], eh[np.minimum(col & rm, len(eh) - 1)], 0)      # lanes right of the window read nothing
        nib = q[np.minimum(col, qlen + 63)]
        if generic:
            sc = mat25[5 * tb + np.minimum(nib, 4)]
            Wm = (wd + (sc << 16)) & 0xffffffff
        else:
            match = (nib == tb) & (K[None, :] >= np.clip(lo, 0, 8)[:, None])
            Wm = (wd + np.where(match, (a + b) << 16, 0)) & 0xffffffff
        Wsh = (Wm << 16) & 0xffffffff
        if variant == 2:
            mkp = add_max2(Wm, c_mis, 0)
            guard = (0 - (wd & 0xffff0000)) & 0xffffffff
            mk = np.minimum(mkp, guard)
            hh = max2(mk, Wsh); g = add_max2(mk, c_noe_ins, 0); t = add_max2(mk, c_noe_del, 0)
        else:
            hh = add_max2(Wm, c_mis, Wsh); g = add_max2(hh, c_noe_ins, 0); t = None
        fl = np.zeros((32, cpl), dtype=np.int64); run = np.zeros(32, dtype=np.int64)
        for k in range(cpl):
            if variant == 1 and generic: run = np.where(lo == k, 0, run)
            fl[:, k] = run
            run = add_max2(run, c_ne_ins, g[:, k])
        eC = cpl * e_ins
        runi = run >> 16
        P = runi + eC * L32
        d = 1
        while d < 32:
            o = np.concatenate([np.zeros(d, dtype=np.int64), P[:-d]])
            P = np.where(L32 >= d, np.maximum(P, o), P); d <<= 1
        Pex = np.concatenate([[0], P[:-1]])
        fin0 = np.where(L32 > 0, np.maximum(Pex - eC * (L32 - 1), 0), 0)
        cin = max(carry, 0)
        uin = np.maximum(fin0, cin - eC * L32)
        carry = int(max(runi[31], uin[31] - eC))
        upk = pack(np.maximum(uin, -1), 0)
        f = max2(fl, (upk[:, None] + K[None, :] * c_ne_ins) & 0xffffffff)
        hq = max2(hh, f)
        if variant == 1: t = add_max2(hq, c_noe_del, 0)
        cand = s32(hq) + K[None, :]                      # (h << 16) + k: ties go to the right
        if generic and variant == 1:                     # looked-up scores: a zeroed column left of j0 can hold max(s, 0) > 0
            cand = np.where(K[None, :] >= lo[:, None], cand, -1)
        lkey = np.max(cand, axis=1)
        key = np.where(hi > 0, np.maximum(key, lkey + jl), key)
        zbits = (hq == 0) if variant == 1 else None
        hleft = np.concatenate([[hcarry], hq[:-1, cpl - 1]])
        hcarry = int(hq[31, cpl - 1])
        hprev = np.concatenate([hleft[:, None], hq[:, :-1]], axis=1)
        ow = add_max2(wd, c_eh, (hprev & 0xffff0000) | (t >> 16))
        touch = (hi >= 0) & (lo < cpl)
        for l in np.nonzero(touch)[0]:
            full = lo[l] < 0 and hi[l] >= cpl
            if variant == 2 and not full:
                for k in range(cpl):
                    if lo[l] <= k <= hi[l]:
                        v = int(ow[l, k])
                        if k == lo[l]: v = (v & 0xffff) | (fc << 16)
                        if k == hi[l]: v &= 0xffff0000
                        eh[(jl[l] + k) & rm] = v; zb[(jl[l] + k) & rm] = int(v != 0)
            else:
                eh[(jl[l] + K) & rm] = ow[l]
                if variant == 1:
                    if lo[l] > 0: eh[j0 & rm] = (eh[j0 & rm] & 0xffff) | (fc << 16)
                    if hi[l] < cpl: eh[lim & rm] &= 0xffff0000
                    if not single: zb[(jl[l] + K) & rm] = zbits[l].astype(np.int64)
                else:
                    zb[(jl[l] + K) & rm] = (ow[l] != 0).astype(np.int64)
        return carry, hcarry, key, (zbits, jl, cpl)

    def narrow_row(cpl, j0, lim, fc, tb):
        """k2_narrow_row: one round, lane l owns the cpl columns j0 + cpl*l + k (no alignment), scalar cell, live masks."""
        K = np.arange(cpl, dtype=np.int64)
        c0 = j0 + cpl * L32
        col = c0[:, None] + K[None, :]
        live = col < lim
        wd = np.where(live, eh[np.minimum(col & rm, len(eh) - 1)], 0)
        nib = q[np.minimum(col, qlen + 63)]
        sc = mat25[5 * tb + np.minimum(nib, 4)] if generic else np.where(nib == tb, a, -b)
        sc = np.where(live | generic, sc, np.where(0 == tb, a, -b))          # dead columns read a zero query word
        M, e = wd >> 16, wd & 0xffff
        hh = np.maximum(M + sc, e)
        g = np.where(live, np.maximum(hh - oe_ins, 0), 0)
        fl = np.zeros((32, cpl), dtype=np.int64); run = np.zeros(32, dtype=np.int64)
        for k in range(cpl):
            fl[:, k] = run
            run = np.maximum(run - e_ins, g[:, k])
        estep = cpl * e_ins
        P = run + estep * L32
        d = 1
        while d < 32:
            o = np.concatenate([np.zeros(d, dtype=np.int64), P[:-d]])
            P = np.where(L32 >= d, np.maximum(P, o), P); d <<= 1
        Pex = np.concatenate([[0], P[:-1]])
        u = np.where(L32 > 0, np.maximum(Pex - estep * (L32 - 1), 0), 0)
        f = np.maximum(fl, u[:, None] - K[None, :] * e_ins)
        h = np.maximum(hh, f)
        t = np.maximum(h - oe_del, 0)
        enew = np.maximum(s16(e - e_del), t)                                  # 16-bit half of the packed add-max
        key = int(np.where(live, h * 65536 + col, -1).max())
        hleft = np.concatenate([[fc], h[:-1, cpl - 1]])
        hprev = np.concatenate([hleft[:, None], h[:, :-1]], axis=1)
        hprev = np.where(col == j0, fc, hprev)
        for l in range(32):
            for k in range(cpl):
                c = int(col[l, k])
                if c <= lim: eh[c & rm] = pack(int(hprev[l, k]), int(enew[l, k]) if c < lim else 0)
        mj = key & 0xffff
        hlast = int(h[(lim - 1 - j0) // cpl, (lim - 1 - j0) % cpl])
        z = live & (h == 0)
        za = z & (col <= mj - 1); ze = z & (col >= mj + 1)
        return key, hlast, (int(col[za].max()) if za.any() else -1), (int(col[ze].min()) if ze.any() else 0x7fffffff)

    for i in range(tlen):
        tb = int(target[i])
        j0 = max(beg, i - w); lim = min(end, i + w + 1, qlen)
        fc = max(h0 - (o_del + e_del * (i + 1)), 0) if (variant == 1 or j0 == 0) else 0
        if variant == 2 and lim - 1 > hiw:
            for c in range(hiw + 1, lim):
                eh[c & rm] = pack(0 if c > qlen else max(h0 - o_ins - c * e_ins, 0), 0)
        if variant == 2: hiw = max(hiw, lim)
        if lim <= j0:
            if j0 == qlen and not gscore > fc: max_ie, gscore = i, fc
            break
        if narrow and variant == 1 and lim - j0 < 64:
            kmax, h1, cb, ce = narrow_row(1 if lim - j0 < 32 else 2, j0, lim, fc, tb)
            m, mj = kmax >> 16, kmax & 0xffff
            cells += lim - j0
            if lim == qlen and not gscore > h1: max_ie, gscore = i, h1
            if m == 0: break
            if m > mx:
                mx, max_i, max_j = m, i, mj
                max_off = max(max_off, abs(mj - i))
            elif zdrop > 0:
                di, dj = i - max_i, mj - max_j
                if di > dj:
                    if mx - m - (di - dj) * e_del > zdrop: break
                elif mx - m - (dj - di) * e_ins > zdrop: break
            beg = cb + 2 if cb >= 0 else (j0 + 1 if fc == 0 else j0)
            end = ce + 1 if ce != 0x7fffffff else lim + 1
            continue
        span = lim - (j0 & ~7)
        wide12 = variant == 1 and 256 <= span < 384
        single = variant == 1 and (span < 256 or wide12)
        key = np.full(32, -1, dtype=np.int64); carry, hcarry = 0, pack(fc, 0)
        keyprev = key.copy()
        if wide12:
            carry, hcarry, key, last = round_packed(j0 & ~7, 12, j0, lim, fc, tb, carry, hcarry, key, True)
        else:
            rbase = j0 & ~7
            while rbase <= lim:
                keyprev = key.copy()
                carry, hcarry, key, last = round_packed(rbase, 8, j0, lim, fc, tb, carry, hcarry, key, single)
                rbase += 256
        kmax = int(key.max())
        if (kmax & 0xffff) >= lim:                       # a column right of the window won: repair from the row buffer
            zb_, jl_, cpl_ = last
            mk2 = keyprev.copy()
            for l in range(32):
                for k in range(cpl_):
                    c = int(jl_[l]) + k
                    if j0 <= c < lim: mk2[l] = max(mk2[l], int(eh[(c + 1) & rm] & 0xffff0000) + c)
            kmax = int(mk2.max())
        h1 = int(eh[lim & rm] >> 16)
        m, mj = kmax >> 16, kmax & 0xffff
        if single:
            zbits, jl, cpl = last
            colz = jl[:, None] + np.arange(cpl)[None, :]
            za = zbits & (colz >= j0) & (colz <= mj - 1); ze = zbits & (colz >= mj + 1) & (colz <= lim - 1)
            cb = int(colz[za].max()) if za.any() else -1
            ce = int(colz[ze].min()) if ze.any() else 0x7fffffff
        elif variant == 2:
            nzc = [c for c in range(j0, lim) if zb[c & rm]]
            nbeg = nzc[0] if nzc else lim
            nzl = [c for c in range(nbeg, lim + 1) if zb[c & rm]]
            cb, ce = nbeg, min((nzl[-1] if nzl else nbeg - 1) + 2, qlen)
        else:
            za = [c for c in range(j0, mj) if zb[c & rm]]; ze = [c for c in range(mj + 1, lim) if zb[c & rm]]
            cb = za[-1] if za else -1
            ce = ze[0] if ze else 0x7fffffff
        cells += lim - j0
        if lim == qlen and not gscore > h1: max_ie, gscore = i, h1
        if m == 0: break
        if m > mx:
            mx, max_i, max_j = m, i, mj
            max_off = max(max_off, abs(mj - i))
        elif zdrop > 0:
            di, dj = i - max_i, mj - max_j
            if di > dj:
                if mx - m - (di - dj) * e_del > zdrop: break
            elif mx - m - (dj - di) * e_ins > zdrop: break
        if variant == 2: beg, end = cb, ce
        else:
            beg = cb + 2 if cb >= 0 else (j0 + 1 if fc == 0 else j0)
            end = ce + 1 if ce != 0x7fffffff else lim + 1
    return mx, max_j + 1, max_i + 1, max_ie + 1, gscore, max_off, cells


def s32(x):
    x = np.asarray(x, dtype=np.int64) & 0xffffffff
    return (x ^ 0x80000000) - 0x80000000


def check(O, tasks, variant, **pk):
    p = O.make_params(**pk)
    ro, co = O.extend_batch(p, tasks["qbuf"], tasks["qoff"], tasks["tbuf"], tasks["toff"], tasks["h0"], tasks["w"], variant=variant)
    for i in range(len(tasks["h0"])):
        qs = tasks["qbuf"][tasks["qoff"][i]:tasks["qoff"][i + 1]]; ts = tasks["tbuf"][tasks["toff"][i]:tasks["toff"][i + 1]]
        w = int(O.lib().bswref_clamp_w(C.byref(p), len(qs), int(tasks["w"][i]), p.end_bonus))
        want = tuple(int(ro[k][i]) for k in ("score", "qle", "tle", "gtle", "gscore", "max_off")) + (int(co[i]),)
        for narrow in ((False, True) if variant == 1 else (False,)):          # option k2_narrow = 0 / 1 (the default)
            got = k2_model(p, qs, ts, int(tasks["h0"][i]), w, variant, narrow=narrow)
            assert got == want, (i, variant, narrow, pk, len(qs), len(ts), int(tasks["h0"][i]), w, got, want)


def near_matches(rng, n, qlo, qhi, h0, w, sub=0.02, tail=40):
    qs, ts = [], []
    for _ in range(n):
        ql = int(rng.integers(qlo, qhi))
        q = rng.integers(0, 4, ql).astype(np.uint8)
        t = q.copy(); t[rng.random(ql) < sub] = rng.integers(0, 4)
        cut = int(rng.integers(ql // 3, ql // 2))
        t = np.concatenate([t[:cut], rng.integers(0, 4, int(rng.integers(0, 6))).astype(np.uint8), t[cut:], rng.integers(0, 4, tail).astype(np.uint8)])
        qs.append(q); ts.append(t)
    qbuf, qoff, tbuf, toff = flat_from_lists(qs, ts)
    return dict(qbuf=qbuf, qoff=qoff, tbuf=tbuf, toff=toff, h0=np.full(n, h0, np.int32), w=np.full(n, w, np.int32))


@pytest.mark.parametrize("variant", [1, 2])
def test_small_tasks(O, variant):
    rng = np.random.default_rng(500 + variant)
    check(O, random_small_tasks(rng, 120, qmax=90, tmax=120), variant)
    t = random_small_tasks(rng, 80, qmax=140, tmax=200)
    t["h0"] = rng.integers(1, 200, len(t["h0"])).astype(np.int32)
    m = rng.random(len(t["qbuf"])) < 0.01; t["qbuf"] = np.where(m, 4, t["qbuf"]).astype(np.uint8)      # N: matrix lookup
    check(O, t, variant, o_del=4, e_del=2, o_ins=7, e_ins=1, zdrop=30, a=2, b=3)


@pytest.mark.parametrize("variant", [1, 2])
def test_wide_windows_take_several_rounds_or_twelve_columns(O, variant):
    """Windows of 250-400 columns: 12-column rounds (V1), two 8-column rounds, shared-memory narrowing."""
    rng = np.random.default_rng(510 + variant)
    check(O, near_matches(rng, 3, 500, 700, 300, 170), variant, zdrop=0)
    check(O, near_matches(rng, 2, 500, 700, 400, 250), variant)


def test_rows_that_die_next_to_a_matching_dead_column(O):
    """Matrix-lookup scoring (an N in the query), V1, tiny h0: the task ends on a row whose live cells are all zero while a
    zeroed column left of the window matches the target base (h = +a there).  Such a column must not win the arg-max --
    found with this model in round 2 (the kernel's registers-only path for rows below 64 columns had hidden it)."""
    rng = np.random.default_rng(530)
    for rep in range(2):
        t = random_small_tasks(rng, 150, qmax=60, tmax=90)
        m = rng.random(len(t["qbuf"])) < 0.03; t["qbuf"] = np.where(m, 4, t["qbuf"]).astype(np.uint8)
        t["h0"] = rng.integers(1, 6, len(t["h0"])).astype(np.int32)
        check(O, t, 1, a=int(rng.integers(1, 4)), b=int(rng.integers(1, 5)), zdrop=0)


def test_ring_row_buffer_and_the_16_bit_cap(O):
    """A query beyond 2 048 columns runs on the ring (lanes of a 12-column round straddle its end); h0 = 32767 - qlen
    drives the packed cell to the cap, where the add inside the instruction wraps and the maximum brings it back."""
    rng = np.random.default_rng(520)
    t = near_matches(rng, 1, 2300, 2400, 1, 150, sub=0.004)
    t["h0"][:] = 32767 - (t["qoff"][1] - t["qoff"][0])
    check(O, t, 1, zdrop=0)
    check(O, t, 2, zdrop=100)
