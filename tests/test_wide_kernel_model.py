"""CPU model of K5's algorithm (bwa-mem-sw_b200/csrc/bsw_k5.cu): the 32-bit extension kernel processes a row 32 columns
at a time, solves the F chain as a max-plus prefix scan with a carry between the chunks, finds the row's arg-max with a
"last lane that holds the chunk maximum" rule and narrows the band with ballots over the row buffer.  This file restates
those steps lane-parallel in numpy (one array element per lane, the same chunk loop, the same slot semantics of the row
buffer) and checks them against the scalar oracle -- so the ALGORITHM is covered by the CPU suite; the kernel itself is
compared with the oracle on the GPU (tests/test_gpu_parity.py: test_every_task_on_the_32_bit_kernel,
test_tasks_beyond_16_bits_in_a_mixed_batch)."""
import ctypes as C

import numpy as np
import pytest

from helpers import random_small_tasks

NEG = -0x3fffffff
LANES = np.arange(32)


def wide_model(p, query, target, h0, w, variant):
    """One extension the way k5_task runs it.  Returns (score, qle, tle, gtle, gscore, max_off, cells)."""
    mat = np.frombuffer(bytes(p.mat), dtype=np.int8).astype(np.int64)
    o_del, e_del, o_ins, e_ins, zdrop = p.o_del, p.e_del, p.o_ins, p.e_ins, p.zdrop
    oe_del, oe_ins = o_del + e_del, o_ins + e_ins
    qlen, tlen = len(query), len(target)
    q = np.asarray(query, dtype=np.int64)
    H = np.zeros(qlen + 2, dtype=np.int64); E = np.zeros(qlen + 2, dtype=np.int64)
    H[0] = h0
    for j in range(1, qlen + 1):
        H[j] = max(h0 - oe_ins - (j - 1) * e_ins, 0)
    max_sc, max_i, max_j, max_ie, gscore, max_off = h0, -1, -1, -1, -1, 0
    beg, end, ncell = 0, qlen, 0
    for i in range(tlen):
        srow = mat[5 * int(target[i]): 5 * int(target[i]) + 5]
        beg = max(beg, i - w); end = min(end, i + w + 1, qlen)
        h1 = max(h0 - (o_del + e_del * (i + 1)), 0) if (variant == 1 or beg == 0) else 0
        m, mj, hcarry, fcarry = 0, -1, h1, 0
        c = beg
        while c < end:
            j = c + LANES
            on = j < end
            jj = np.where(on, j, 0)
            M = np.where(on, H[jj], 0); e = np.where(on, E[jj], 0); sc = np.where(on, srow[q[np.minimum(jj, qlen - 1)]], 0)
            M = M + sc if variant == 1 else np.where(M != 0, M + sc, 0)
            x = np.maximum(M, e)
            gk = np.maximum((x if variant == 1 else M) - oe_ins, 0)
            u = np.where(on, gk + j * e_ins, NEG)
            d = 1
            while d < 32:                                           # inclusive max scan (shfl_up)
                o = np.concatenate([np.full(d, NEG), u[:-d]])
                u = np.where(LANES >= d, np.maximum(u, o), u)
                d <<= 1
            ux = np.concatenate([[NEG], u[:-1]])
            f = fcarry - LANES * e_ins
            f = np.where(LANES > 0, np.maximum(f, ux - (j - 1) * e_ins), f)
            h = np.maximum(x, f)
            fcarry = max(fcarry - 32 * e_ins, int(u[31]) - (c + 31) * e_ins)
            t = np.maximum((h if variant == 1 else M) - oe_del, 0)
            e = np.maximum(e - e_del, t)
            hl = np.concatenate([[hcarry], h[:-1]])
            H[jj[on]] = hl[on]; E[jj[on]] = e[on]
            hcarry = int(h[min(31, end - 1 - c)])
            cm = int(np.where(on, h, -1).max())
            if cm >= m:
                hit = np.nonzero(on & (h == cm))[0]
                m, mj = cm, c + int(hit[-1])
            c += 32
        H[end] = hcarry; E[end] = 0
        if end > beg: ncell += end - beg
        if (end if end > beg else beg) == qlen:
            if not gscore > hcarry: max_ie, gscore = i, hcarry
        if m == 0: break
        if m > max_sc:
            max_sc, max_i, max_j = m, i, mj
            max_off = max(max_off, abs(mj - i))
        elif zdrop > 0:
            di, dj = i - max_i, mj - max_j
            if di > dj:
                if max_sc - m - (di - dj) * e_del > zdrop: break
            elif max_sc - m - (dj - di) * e_ins > zdrop: break
        if variant == 1:                                            # ballots over the row buffer, 32 slots at a time
            nb, c = beg, mj
            while c >= beg:
                j = c - LANES
                z = (j >= beg) & (H[np.maximum(j, 0)] == 0)
                if z.any(): nb = c - int(np.nonzero(z)[0][0]) + 1; break
                c -= 32
            ne, c = end + 1, mj + 2
            while c <= end:
                j = c + LANES
                z = (j <= end) & (H[np.minimum(j, qlen + 1)] == 0)
                if z.any(): ne = c + int(np.nonzero(z)[0][0]); break
                c += 32
            if mj + 2 > end: ne = mj + 2
            beg, end = nb, ne
        else:
            nb, c = end, beg
            while c < end:
                j = c + LANES
                jc = np.minimum(j, qlen + 1)
                nz = (j < end) & ((H[jc] != 0) | (E[jc] != 0))
                if nz.any(): nb = c + int(np.nonzero(nz)[0][0]); break
                c += 32
            jl, c = nb - 1, end
            while c >= nb:
                j = c - LANES
                jc = np.maximum(j, 0)
                nz = (j >= nb) & ((H[jc] != 0) | (E[jc] != 0))
                if nz.any(): jl = c - int(np.nonzero(nz)[0][0]); break
                c -= 32
            beg, end = nb, min(jl + 2, qlen)
    return max_sc, max_j + 1, max_i + 1, max_ie + 1, gscore, max_off, ncell


def check(O, tasks, variant, **pk):
    p = O.make_params(**pk)
    ro, co = O.extend_batch(p, tasks["qbuf"], tasks["qoff"], tasks["tbuf"], tasks["toff"], tasks["h0"], tasks["w"], variant=variant)
    n = len(tasks["h0"])
    for i in range(n):
        qs = tasks["qbuf"][tasks["qoff"][i]:tasks["qoff"][i + 1]]; ts = tasks["tbuf"][tasks["toff"][i]:tasks["toff"][i + 1]]
        w = int(O.lib().bswref_clamp_w(C.byref(p), len(qs), int(tasks["w"][i]), p.end_bonus))     # ksw_extend2's band clamp (the host does it)
        got = wide_model(p, qs, ts, int(tasks["h0"][i]), w, variant)
        want = tuple(int(ro[k][i]) for k in ("score", "qle", "tle", "gtle", "gscore", "max_off")) + (int(co[i]),)
        assert got == want, (i, variant, pk, got, want)


@pytest.mark.parametrize("variant", [1, 2])
def test_scan_and_narrowing_model_equals_oracle(O, variant):
    rng = np.random.default_rng(300 + variant)
    check(O, random_small_tasks(rng, 250, qmax=90, tmax=120), variant)
    t = random_small_tasks(rng, 150, qmax=140, tmax=200)
    t["h0"] = rng.integers(1, 200, len(t["h0"])).astype(np.int32)                    # wide live windows: several chunks per row
    check(O, t, variant, o_del=4, e_del=2, o_ins=7, e_ins=1, zdrop=30, a=2, b=3)
    check(O, t, variant, zdrop=0)


def test_scores_beyond_16_bits(O):
    """The model (and the kernel) keep 32-bit rows: h0 far above 32767 on a few hundred columns."""
    rng = np.random.default_rng(310)
    t = random_small_tasks(rng, 60, qmax=200, tmax=260)
    t["h0"] = rng.integers(40_000, 2_000_000, len(t["h0"])).astype(np.int32)
    for variant in (1, 2):
        check(O, t, variant, zdrop=0)
        check(O, t, variant)
