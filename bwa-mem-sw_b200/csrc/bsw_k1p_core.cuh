// bsw_k1p_core.cuh -- K1P: two extension tasks per lane, scores packed two-per-register (int16x2 SIMD-in-word).
//
// Same recurrence, same lazy narrowing and the same bit-exact outputs as bsw_k1_core.cuh (one FPGA PE's sw_extend,
// sw_pe_array_sw_extend.v FSM :1639-1705), but a lane carries TWO tasks A and B of similar shape (the scheduler pairs
// neighbours of the length-sorted order): every DPX instruction of the branch-free cell now advances two DP cells.
//
// Row state per column per lane: two words  HW = {H_B[31:16], H_A[15:0]}  EW = {E_B, E_A}  (uint2 at eh[(j*32+lane)*2],
// one LDS.64 / STS.64 per pair of cells).  The query of each task is re-coded into one-hot planes (positions of A, C, G, T
// per 32 columns) exactly as in K1.  Per pair of cells, branch-free path:
//     v   = (c >> k) & 0x00010001            match bits of A (bit 0) and B (bit 16)            SHF + LOP3
//     HWm = v * (a+b) + HW                   {M_B + (a+b)*match_B, M_A + (a+b)*match_A}        IMAD
//     hh  = max(HWm + {-b,-b}, EW)           sx:1797,1798                                       VIADDMNMX.S16x2
//     g   = max(hh + {-oe_ins}, 0)           sx:1863,1865 (f - oe_ins <= f - e_ins, so hh suffices: 1-op F chain)
//     h   = max(hh, f)                       sx:1809                                            VIMNMX.S16x2
//     t   = max(h + {-oe_del}, 0)            sx:1866,1862
//     EW' = max(EW + {-e_del}, t)            sx:1770-1771
//     f   = max(f + {-e_ins}, g)             sx:1780-1781
//     keyA = max(keyA, (h << 16) + k) ; keyB = max(keyB, (h & 0xffff0000) + k)                  sx:1808,1816
//     store {h1, EW'} ; h1 = h               sx:1776
//
// The two tasks of a lane have their own windows [j0, lim).  What a task's half holds outside its own window is
// irrelevant (left of beg is never read again, right of the end slot is rewritten before it is read), so:
//   * columns where only one task has cells (head / tail of the union) run that task alone, cell by cell, on its
//     16-bit halves (this path also handles every narrowing event);
//   * the common part runs packed; a finished task's half just computes garbage next to its live partner.
#pragma once
#include "bsw_k1_core.cuh"

namespace bsw {

constexpr int K1P_CS = 2 * TILE_LANES;      // words between consecutive columns of one lane (uint2 per lane)

struct K1PTask {
    // task constants
    int qlen, tlen, h0, w;
    // running results (sx:889,1009,919,1019,1029,929) and narrowing state
    int max, max_i, max_j, max_ie, gscore, max_off;
    int beg, cend, resetmax, stopmin;
    uint32_t cells;
    bool done;
    // target stream
    uint32_t tw, tnext;
    // row scope
    bool inrow, stopped;
    uint32_t tb;
    int j0, lim, fc, beff;
    int f, h1, key;
};

BSW_HD void k1p_init(K1PTask& T, const SlotParam& sp, const uint32_t* tg)
{
    T.qlen = sp.qlen; T.tlen = sp.tlen; T.h0 = sp.h0; T.w = sp.w;
    T.max = sp.h0; T.max_i = -1; T.max_j = -1; T.max_ie = -1; T.gscore = -1; T.max_off = 0;
    T.beg = 0; T.cend = sp.qlen; T.resetmax = -1; T.stopmin = 0x7fffffff;
    T.cells = 0; T.done = sp.qlen <= 0;
    T.tw = 0; T.tnext = 0;
    if (!T.done) { T.tw = tg[0]; if (sp.tlen > 8) T.tnext = tg[K1_S]; }
    T.inrow = false; T.stopped = false; T.tb = 0; T.j0 = 0; T.lim = 0; T.fc = 0; T.beff = 0; T.f = 0; T.h1 = 0; T.key = K1_KEY_NONE;
}

// One-hot planes of one task's query, in place (same as K1).
BSW_HD void k1p_planes(uint32_t* qs, int qlen, int nqw_tile)
{
    const int nblk = qlen > 0 ? ((qlen + 31) >> 5) + 1 : 0;
    for (int m = 0; m < nblk; ++m) {
        uint32_t wd[4];
        for (int u = 0; u < 4; ++u) wd[u] = (4 * m + u < nqw_tile) ? qs[(4 * m + u) * K1_S] : 0u;
        for (uint32_t b = 0; b < 4; ++b) {
            uint32_t pl = k1_eq8(wd[0], b) | (k1_eq8(wd[1], b) << 8) | (k1_eq8(wd[2], b) << 16) | (k1_eq8(wd[3], b) << 24);
            if (32 * m >= qlen) pl = 0;
            qs[(4 * m + (int)b) * K1_S] = pl;
        }
    }
}

// 16-bit views of one task's half of the row buffer.  eh points at this lane's uint2 of column 0.
template <int X> BSW_HD int k1p_ldH(const uint32_t* eh, int j) { return (int)reinterpret_cast<const uint16_t*>(eh + j * K1P_CS)[X]; }
template <int X> BSW_HD int k1p_ldE(const uint32_t* eh, int j) { return (int)reinterpret_cast<const uint16_t*>(eh + j * K1P_CS + 1)[X]; }
template <int X> BSW_HD void k1p_stH(uint32_t* eh, int j, int v) { reinterpret_cast<uint16_t*>(eh + j * K1P_CS)[X] = (uint16_t)v; }
template <int X> BSW_HD void k1p_stE(uint32_t* eh, int j, int v) { reinterpret_cast<uint16_t*>(eh + j * K1P_CS + 1)[X] = (uint16_t)v; }

// Row prologue of one task: target base, window, first column, lazy-narrowing trims (identical to k1_task).
template <int X>
BSW_HD void k1p_row_begin(K1PTask& T, const DevParams& P, int i, const uint32_t* eh, const uint32_t* tg)
{
    T.inrow = !T.done && i < T.tlen;
    T.stopped = false;
    if (!T.inrow) { if (!T.done) T.done = true; return; }
    if ((i & 7) == 0 && i) {
        T.tw = T.tnext;
        if (i + 8 < T.tlen) T.tnext = tg[((i >> 3) + 1) * K1_S];
    }
    T.tb = T.tw & 15u;
    T.tw >>= 4;
    int j0 = imax(T.beg, i - T.w);                                               // sx:1846,1894,1895,1803
    int lim = imin(imin(T.cend, i + T.w + 1), T.qlen);                           // sx:1980,1843,1897,1898,1842
    if (T.stopmin < j0) {
        const int zend = imin(j0, lim);
        for (int z = T.stopmin; z < zend; ++z)
            if (k1p_ldH<X>(eh, z) == 0) { lim = imin(lim, z); break; }
    }
    T.fc = imax(T.h0 - (P.o_del + P.e_del * (i + 1)), 0);                        // V1: unconditional (sx:1796,1795,1880,1835,849)
    while (j0 < lim && j0 <= T.resetmax && k1p_ldH<X>(eh, j0) == 0) ++j0;       // beg' = last zero + 1 (sx:1766-1769)
    while (lim > j0 && lim - 1 >= T.stopmin && k1p_ldH<X>(eh, lim - 1) == 0) --lim;   // end' = first zero >= mj+2 (sx:1779,1782-1789)
    T.j0 = j0; T.lim = lim; T.beff = j0;
    T.f = 0; T.h1 = T.fc; T.key = K1_KEY_NONE;
}

// Cells [from, to) of one task alone, cell by cell, with the narrowing events.  A "stop" event shrinks T.lim.
template <int X>
BSW_HD void k1p_single(K1PTask& T, const DevParams& P, int from, int to, uint32_t* eh, const uint32_t* qs)
{
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins, e_del = P.e_del, e_ins = P.e_ins;
    const int mat = P.match, mis = -P.mismatch;
    const uint32_t* plane = qs + (T.tb & 3u) * K1_S;
    int f = T.f, h1 = T.h1, key = T.key;
    for (int j = from; j < to; ++j) {
        const int M = k1p_ldH<X>(eh, j);
        int e = k1p_ldE<X>(eh, j);
        if (M == 0) {
            if (j <= T.resetmax) { f = 0; h1 = T.fc; key = K1_KEY_NONE; T.beff = j + 1; continue; }      // beg' = j+1
            if (j >= T.stopmin) { T.lim = j; T.stopped = true; break; }                                 // end' = j
        }
        const int s = ((plane[(4 * (j >> 5)) * K1_S] >> (j & 31)) & 1u) ? mat : mis;
        const int h = imax(imax(M + s, e), f);                                   // sx:1797,1798,1809
        key = imax(key, h * 65536 + j);                                          // sx:1808,1816
        int t = imax(h - oe_del, 0);                                             // sx:1866,1862
        e = imax(e - e_del, t);                                                  // sx:1770-1771
        t = imax(h - oe_ins, 0);                                                 // sx:1863,1865
        f = imax(f - e_ins, t);                                                  // sx:1780-1781
        k1p_stH<X>(eh, j, h1);                                                   // sx:1776
        k1p_stE<X>(eh, j, e);
        h1 = h;
    }
    T.f = f; T.h1 = h1; T.key = key;
}

// Row epilogue of one task (identical to k1_task).
template <int X>
BSW_HD void k1p_row_end(K1PTask& T, const DevParams& P, int i, uint32_t* eh)
{
    if (!T.inrow) return;
    const int e_eff = T.lim, b_eff = T.beff;
    if (e_eff > b_eff) T.cells += (uint32_t)(e_eff - b_eff);
    k1p_stH<X>(eh, e_eff, T.h1);                                                 // eh[end] = {h1, e=0}: sx:1775,1904
    k1p_stE<X>(eh, e_eff, 0);
    const int j_after = e_eff > b_eff ? e_eff : b_eff;
    if (j_after == T.qlen) {                                                     // sx:1768,1913
        if (!(T.gscore > T.h1)) { T.max_ie = i; T.gscore = T.h1; }               // sx:1941,1829,1831
    }
    int m, mj;
    if (T.key < 0) { m = 0; mj = -1; } else { m = T.key >> 16; mj = T.key & 0xffff; }
    if (m == 0) { T.done = true; return; }                                       // sx:1942
    if (m > T.max) {                                                             // sx:1959
        T.max = m; T.max_i = i; T.max_j = mj;
        const int d = mj > i ? mj - i : i - mj;
        T.max_off = T.max_off > d ? T.max_off : d;                               // sx:1707-1708,1812
    } else if (P.zdrop > 0) {                                                    // ksw_extend2 z-drop (not in the RTL)
        const int di = i - T.max_i, dj = mj - T.max_j;
        if (di > dj) { if (T.max - m - (di - dj) * P.e_del > P.zdrop) { T.done = true; return; } }
        else         { if (T.max - m - (dj - di) * P.e_ins > P.zdrop) { T.done = true; return; } }
    }
    T.beg = b_eff; T.cend = e_eff + 1; T.resetmax = mj; T.stopmin = mj + 2;      // lazy form of sx:1766-1769,1779,1782-1789
    if (i + 1 >= T.tlen) T.done = true;
}

BSW_HD void k1p_result(const K1PTask& T, SlotResult& r)
{
    r.score = T.max; r.qle = T.max_j + 1; r.tle = T.max_i + 1; r.gtle = T.max_ie + 1;     // sx:1315-1375,1841,1868,1794
    r.gscore = T.gscore; r.max_off = T.max_off; r.cells = (int32_t)T.cells; r.status = STATUS_OK;
}

// Two extensions per lane.  eh: this lane's uint2 of column 0 (column stride K1P_CS words); qsA/qsB: this lane's packed
// query words of the two tasks (stride K1_S, room for nqw_max + K1_QS_EXTRA words); tgA/tgB: target words (stride K1_S).
template <int SYM>
BSW_HD void k1p_pair(const DevParams& P, const SlotParam& spA, const SlotParam& spB, int nqwA, int nqwB,
                     uint32_t* eh, uint32_t* qsA, uint32_t* qsB, const uint32_t* tgA, const uint32_t* tgB,
                     SlotResult& resA, SlotResult& resB)
{
    const int e_ins = P.e_ins;
    const int mis = -P.mismatch;
    // packed constants, kept in registers
    uint32_t c_ab = (uint32_t)(P.match + P.mismatch);
    uint32_t c_mis = ((uint32_t)mis << 16) | ((uint32_t)mis & 0xffffu);
    uint32_t c_noe_del = ((uint32_t)(-(P.o_del + P.e_del)) << 16) | ((uint32_t)(-(P.o_del + P.e_del)) & 0xffffu);
    uint32_t c_noe_ins = ((uint32_t)(-(P.o_ins + P.e_ins)) << 16) | ((uint32_t)(-(P.o_ins + P.e_ins)) & 0xffffu);
    uint32_t c_ne_del = ((uint32_t)(-P.e_del) << 16) | ((uint32_t)(-P.e_del) & 0xffffu);
    uint32_t c_ne_ins = ((uint32_t)(-e_ins) << 16) | ((uint32_t)(-e_ins) & 0xffffu);
    uint32_t zero = P.zero;                      // 0, opaque to ptxas
    uint32_t c_bits = 0x00010001u;
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+r"(c_ab), "+r"(c_mis), "+r"(c_noe_del), "+r"(c_noe_ins), "+r"(c_ne_del), "+r"(c_ne_ins), "+r"(c_bits));
#endif

    K1PTask A, B;
    k1p_init(A, spA, tgA);
    k1p_init(B, spB, tgB);
    k1p_planes(qsA, A.qlen, nqwA);
    k1p_planes(qsB, B.qlen, nqwB);

    // first row: eh[j].h = H(-1, j-1), all e = 0 (sx:1818; 1979,1957,1974; 1975-1978,1821)
    {
        const int qm = imax(A.qlen, B.qlen);
        int ha = A.h0 - P.o_ins, hb = B.h0 - P.o_ins;
        eh[0] = ((uint32_t)imax(B.h0, 0) << 16) | (uint32_t)imax(A.h0, 0);
        eh[1] = 0;
        for (int j = 1; j <= qm; ++j) {
            ha -= e_ins; hb -= e_ins;
            eh[j * K1P_CS] = ((uint32_t)imax(hb, 0) << 16) | (uint32_t)imax(ha, 0);
            eh[j * K1P_CS + 1] = 0;
        }
    }

    const int tmax = imax(A.done ? 0 : A.tlen, B.done ? 0 : B.tlen);
    for (int i = 0; i < tmax; ++i) {                                             // sx:1891
        k1p_row_begin<0>(A, P, i, eh, tgA);
        k1p_row_begin<1>(B, P, i, eh, tgB);
        if (!A.inrow && !B.inrow) break;

        // the part of the row where the packed path runs: the common window, or the whole window of a lone task
        int cs, ce;
        if (A.inrow && B.inrow) {
            cs = imax(A.j0, B.j0); ce = imin(A.lim, B.lim);
            if (A.j0 < cs) k1p_single<0>(A, P, A.j0, imin(cs, A.lim), eh, qsA);  // head of the union: one task alone
            if (B.j0 < cs) k1p_single<1>(B, P, B.j0, imin(cs, B.lim), eh, qsB);
            if (A.stopped || B.stopped) ce = cs;                                 // a stop event in the head: no packed part
        } else if (A.inrow) { cs = A.j0; ce = A.lim; }
        else { cs = B.j0; ce = B.lim; }

        int j = cs;
        if (j < ce) {
            const bool la = A.inrow, lb = B.inrow;
            const uint32_t* planeA = qsA + (A.tb & 3u) * K1_S;
            const uint32_t* planeB = qsB + (B.tb & 3u) * K1_S;
            uint32_t f = ((uint32_t)B.f << 16) | (uint32_t)A.f;                  // packed {B, A}
            uint32_t h1 = ((uint32_t)B.h1 << 16) | (uint32_t)A.h1;
            int keyA = A.key, keyB = B.key;
            uint32_t* ehp = eh + j * K1P_CS;
            bool event = false;

#define BSW_K1P_CELL(K, HW, EW, LIVE)                                                                \
            {                                                                                        \
                const uint32_t v = ((K) ? (c >> (K)) : c) & c_bits;                                  \
                const uint32_t HWm = v * c_ab + (HW);                                                \
                const uint32_t hh = add_max_s16x2(HWm, c_mis, (EW));                                 \
                const uint32_t h = max_s16x2(hh, f);                                                 \
                const uint32_t t = add_max_s16x2(h, c_noe_del, zero);                                \
                const uint32_t en = add_max_s16x2((EW), c_ne_del, t);                                \
                const uint32_t fn = add_max_s16x2(f, c_ne_ins, add_max_s16x2(hh, SYM ? c_noe_del : c_noe_ins, zero)); \
                if (LIVE) {                                                                          \
                    f = fn;                                                                          \
                    ehp[(K) * K1P_CS] = h1; ehp[(K) * K1P_CS + 1] = en;                              \
                    ckA = (K) ? add_max((int)(h << 16), (K), ckA) : (int)(h << 16);                  \
                    ckB = (K) ? add_max((int)(h & 0xffff0000u), (K), ckB) : (int)(h & 0xffff0000u);  \
                    h1 = h;                                                                          \
                }                                                                                    \
            }

            while (j < ce) {
                const int nv = ce - j;
                const uint32_t w0 = ehp[0 * K1P_CS], w1 = ehp[1 * K1P_CS], w2 = ehp[2 * K1P_CS], w3 = ehp[3 * K1P_CS];
                const uint32_t w4 = ehp[4 * K1P_CS], w5 = ehp[5 * K1P_CS], w6 = ehp[6 * K1P_CS], w7 = ehp[7 * K1P_CS];
                const uint32_t e0 = ehp[0 * K1P_CS + 1], e1 = ehp[1 * K1P_CS + 1], e2 = ehp[2 * K1P_CS + 1], e3 = ehp[3 * K1P_CS + 1];
                const uint32_t e4 = ehp[4 * K1P_CS + 1], e5 = ehp[5 * K1P_CS + 1], e6 = ehp[6 * K1P_CS + 1], e7 = ehp[7 * K1P_CS + 1];
                // a zero H of a live task inside its event region (col <= mj or col >= mj+2) needs the cell-by-cell path
                const uint32_t ones = 0xffffffffu;
                uint32_t zm = min3_u16x2(w0, nv > 1 ? w1 : ones, nv > 2 ? w2 : ones);
                zm = min3_u16x2(zm, nv > 3 ? w3 : ones, nv > 4 ? w4 : ones);
                zm = min3_u16x2(zm, nv > 5 ? w5 : ones, nv > 6 ? w6 : ones);
                zm = min3_u16x2(zm, nv > 7 ? w7 : ones, ones);
                const int jlast = j + imin(nv, 8) - 1;
                if (la && (zm & 0xffffu) == 0 && (j <= A.resetmax || jlast >= A.stopmin)) { event = true; break; }
                if (lb && (zm >> 16) == 0 && (j <= B.resetmax || jlast >= B.stopmin)) { event = true; break; }
                // match bits of columns j..j+7: A in bits 0..7, B in bits 16..23
                const int m = j >> 5, sh = j & 31;
                const uint32_t xa = funnel_r(planeA[(4 * m) * K1_S], planeA[(4 * m + 4) * K1_S], sh);
                const uint32_t xb = funnel_r(planeB[(4 * m) * K1_S], planeB[(4 * m + 4) * K1_S], sh);
                const uint32_t c = (xa & 0xffu) | (xb << 16);
                int ckA = K1_KEY_NONE, ckB = K1_KEY_NONE;
                if (nv >= 8) {
                    BSW_K1P_CELL(0, w0, e0, true)
                    BSW_K1P_CELL(1, w1, e1, true)
                    BSW_K1P_CELL(2, w2, e2, true)
                    BSW_K1P_CELL(3, w3, e3, true)
                    BSW_K1P_CELL(4, w4, e4, true)
                    BSW_K1P_CELL(5, w5, e5, true)
                    BSW_K1P_CELL(6, w6, e6, true)
                    BSW_K1P_CELL(7, w7, e7, true)
                    keyA = imax(keyA, ckA + j); keyB = imax(keyB, ckB + j);
                    j += 8; ehp += 8 * K1P_CS;
                } else {
                    BSW_K1P_CELL(0, w0, e0, true)
                    BSW_K1P_CELL(1, w1, e1, nv > 1)
                    BSW_K1P_CELL(2, w2, e2, nv > 2)
                    BSW_K1P_CELL(3, w3, e3, nv > 3)
                    BSW_K1P_CELL(4, w4, e4, nv > 4)
                    BSW_K1P_CELL(5, w5, e5, nv > 5)
                    BSW_K1P_CELL(6, w6, e6, nv > 6)
                    keyA = imax(keyA, ckA + j); keyB = imax(keyB, ckB + j);
                    j += nv; ehp += nv * K1P_CS;
                }
            }
#undef BSW_K1P_CELL
            // unpack the packed state (a finished task's half is garbage and ignored)
            if (la) { A.f = (int)(f & 0xffffu); A.h1 = (int)(h1 & 0xffffu); A.key = keyA; }
            if (lb) { B.f = (int)(f >> 16); B.h1 = (int)(h1 >> 16); B.key = keyB; }
            (void)event;
        }
        // whatever is left of each task's window (tail of the union, or the rest of the row after an event) runs alone
        if (A.inrow && !A.stopped && j < A.lim) k1p_single<0>(A, P, imax(j, A.j0), A.lim, eh, qsA);
        if (B.inrow && !B.stopped && j < B.lim) k1p_single<1>(B, P, imax(j, B.j0), B.lim, eh, qsB);

        k1p_row_end<0>(A, P, i, eh);
        k1p_row_end<1>(B, P, i, eh);
        if (A.done && B.done) break;
    }
    k1p_result(A, resA);
    k1p_result(B, resB);
}

}  // namespace bsw
