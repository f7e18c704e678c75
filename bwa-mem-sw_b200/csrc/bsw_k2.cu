// bsw_k2.cu -- K2: intra-task extension kernel for long tasks, one CTA of NW = 1 or 4 warps per task (sm_100a).
//
// The reference caps a task at qlen <= 255 / 2048 bases (query_mem 2048x4b, eh_arr 256 entries:
// sw_pe_array_proc_element.v:347-350, sw_pe_array_sw_extend.v:512-515); BASELINE config 4 asks for
// 1-10 kb extensions at w=500, so this kernel has no counterpart in the RTL beyond the recurrence.
//
// Why not an anti-diagonal wavefront: row i+1's window [beg,end) depends on the COMPLETE row i (arg-max
// column mj and the non-zero run around it, sw_pe_array_sw_extend.v:1766-1769,1779,1782-1789), and the
// narrowing is not result-neutral, so a wavefront that starts row i+1 before row i ends cannot be
// bit-exact.  K2 is row-parallel instead: the CTA sweeps one row at a time, warp g taking the g-th 256-column
// group of the window (w = 500 gives a 1001-column band = 4 groups), lane l owning 8 consecutive columns (two
// 128-bit shared-memory accesses each way).  H and E only depend on the previous row.  F is a max-plus linear
// recurrence along the row,
//     f[j+1] = max(f[j] - e_ins, g[j]),   g[j] = max(0, max(M[j]+s[j], e[j]) - oe_ins)
// (g does not need f because f - oe_ins <= f - e_ins for o_ins >= 0), so
//   pass 1: every lane runs its 8 columns with a zero carry-in; the carries are combined across lanes with a
//           5-step __shfl_up_sync prefix-max on A[l] + 8*e_ins*l and across warps through shared memory
//           (A_group - 256*e_ins*distance);
//   pass 2: the carry is folded into f/h and E, the row buffer, the packed arg-max key (REDUX max) and one
//           "H == 0" bit per column are produced.
// The band narrowing is then two masked bit scans over those bits (__clz/__ffs + REDUX), i.e. exactly the
// reference's two scan loops.  Three __syncthreads per row round.
#include <cuda_runtime.h>
#include "bsw_device.cuh"
#include "bsw_k1_core.cuh"
#include "bsw_kernels.h"

namespace bsw {

constexpr int K2_MAXW = 4;               // warps per task: template parameter NW in {1, 4}
constexpr int K2_RING = 2048;            // row-buffer columns kept in shared memory (a ring once the query is longer)
constexpr int K2_GROUP = 256;           // columns per warp step
constexpr int K2_HDR_BYTES = 128;       // mbarrier (8 B) + cross-warp exchange words

__device__ __forceinline__ uint32_t k2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Score of (target row, query nibble): FAST = +match / -mismatch by nibble XOR; GENERIC = byte lookup in the target base's
// matrix row (the RTL's 25:1 mux, sw_pe_array_mux_25to1_sel5_8_1.v:105-142).
template <int GENERIC>
__device__ __forceinline__ int k2_score(uint32_t nib_or_xor, int mat, int mis, uint32_t rlo, uint32_t rhi)
{
    if (GENERIC) return k1_lookup(nib_or_xor, rlo, rhi);
    return nib_or_xor ? mis : mat;
}

// bits [a, b] (inclusive, cell indices) of the 32-cell word that starts at cell `base`
__device__ __forceinline__ uint32_t k2_range_mask(int base, int a, int b)
{
    const int lo = imax(a - base, 0), hi = imin(b - base, 31);
    if (lo > hi) return 0u;
    return (0xffffffffu << lo) & (0xffffffffu >> (31 - hi));
}


// ---- narrow rows (one warp per task) ----
// With PacBio-like errors the narrowed window of a long task is ~50 columns wide (a few hundred at most), so the
// 256-column group of the general path leaves most lanes idle: 5.1 G warp-instructions for 1.3 G cells in round 1, of
// which ~7 lanes per instruction held live cells.  A row whose window is below 32*CPL columns runs here instead: ONE
// round, lane l owns the CPL columns j0 + CPL*l + k (no alignment: the ring is addressed per word), everything stays in
// registers.  The F recurrence is the same zero-carry prefix max, with a stride of CPL columns per lane; the band
// narrowing needs no bit array: one ballot per k gives the zero cells as warp-uniform masks, and the two scans of the
// reference (sx:1766-1769 / 1779,1782-1789) become two find-first/last-set on those masks.
// Outputs: key = (row max << 16 | right-most column holding it), hlast = h of column lim-1, cb / ce = last zero cell
// left of mj / first zero cell right of it (-1 / 0x7fffffff when there is none).
// HI: the row-buffer word is {H hi16, E lo16} (the packed general path below), else {E hi16, H lo16}.
template <int GENERIC, int CPL, bool HI>
__device__ __forceinline__ void k2_narrow_row(uint32_t* __restrict__ eh, const uint32_t* __restrict__ qs, const int rm, const int nqw,
                                              const int j0, const int lim, const int fc, const int lane, const uint32_t trep,
                                              const uint32_t rlo, const uint32_t rhi, const int mat, const int mis,
                                              const int e_ins, const int oe_ins, const int oe_del, const uint32_t ce_pack,
                                              int& key_out, int& hlast, int& cb_out, int& ce_out)
{
    const int c0 = j0 + CPL * lane;
    uint32_t wd[CPL];
    int hh[CPL], fl[CPL], h[CPL];
    bool live[CPL];
    int run = 0;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        const int c = c0 + k;
        live[k] = c < lim;
        wd[k] = live[k] ? eh[c & rm] : 0u;
        const uint32_t qw = (live[k] && (c >> 3) < nqw) ? qs[c >> 3] : 0u;
        const uint32_t nib = (qw >> (4 * (c & 7))) & 15u;
        const int M = HI ? (int)(wd[k] >> 16) : (int)(wd[k] & 0xffffu), e = HI ? (int)(wd[k] & 0xffffu) : (int)(wd[k] >> 16);
        const int sc = k2_score<GENERIC>(GENERIC ? nib : (nib ^ (trep & 15u)), mat, mis, rlo, rhi);
        hh[k] = add_max(M, sc, e);                                                 // sx:1797,1798
        int g = add_max_relu(hh[k], -oe_ins, 0);                                   // sx:1863,1865 with h >= hh
        if (!live[k]) g = 0;
        fl[k] = run;
        run = add_max(run, -e_ins, g);                                             // sx:1780,1781
    }
    const int estep = CPL * e_ins;
    int P = run + estep * lane;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, P, d);
        if (lane >= d) P = imax(P, o);
    }
    const int Pex = __shfl_up_sync(0xffffffffu, P, 1);
    int u = (lane > 0) ? imax(Pex - estep * (lane - 1), 0) : 0;                    // f entering this lane's first column
    int key = -1;
    uint32_t enew[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        const int f = imax(fl[k], u);
        u -= e_ins;
        h[k] = imax(hh[k], f);                                                     // sx:1809
        const int t = add_max_relu(h[k], -oe_del, 0);                              // sx:1866,1862
        enew[k] = add_max_s16x2(wd[k], ce_pack, HI ? (uint32_t)t : ((uint32_t)t << 16));   // E half: max(e-e_del,t), H half: 0  sx:1770-1771
        if (live[k]) key = imax(key, h[k] * 65536 + c0 + k);                       // sx:1808,1816
    }
    int hleft = __shfl_up_sync(0xffffffffu, h[CPL - 1], 1);
    if (lane == 0) hleft = fc;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        const int c = c0 + k;
        if (c <= lim) {                                                            // cells and the end slot eh[lim] = {h1, 0} (sx:1775,1904,1776)
            const int h1 = (c == j0) ? fc : (k ? h[k - 1] : hleft);
            eh[c & rm] = (c < lim ? enew[k] : 0u) | (HI ? ((uint32_t)h1 << 16) : (uint32_t)h1);
        }
    }
    key = __reduce_max_sync(0xffffffffu, key);
    const int mj = key & 0xffff;
    // h of the last cell (column lim-1), for the gscore rule
    {
        const int d = lim - 1 - j0;
        int v = h[0];
        if (CPL == 2) v = (d & 1) ? h[1] : h[0];
        hlast = __shfl_sync(0xffffffffu, v, d / CPL);
    }
    int cb = -1, ce = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        const uint32_t z = __ballot_sync(0xffffffffu, live[k] && h[k] == 0);       // bit l <-> column j0 + CPL*l + k
        const int la = mj - 1 - j0 - k;                                            // lanes l with column <= mj-1: CPL*l <= la
        if (la >= 0) {
            const int lmax = la / CPL;
            const uint32_t m = z & (lmax >= 31 ? 0xffffffffu : ((2u << lmax) - 1u));
            if (m) cb = imax(cb, j0 + CPL * (31 - __clz(m)) + k);
        }
        const int lb = mj + 1 - j0 - k;                                            // lanes l with column >= mj+1: CPL*l >= lb
        const int lmin = lb <= 0 ? 0 : (lb + CPL - 1) / CPL;
        if (lmin < 32) {
            const uint32_t m = z & (0xffffffffu << lmin);
            if (m) ce = imin(ce, j0 + CPL * (__ffs(m) - 1) + k);
        }
    }
    key_out = key; cb_out = cb; ce_out = ce;
}

// ---- the packed round (one warp per task, V1) ----
// K1's cell (bsw_k1_core.cuh): the row-buffer word is {H hi16, E lo16}; H, F and the running maxima live in the high half
// of a register with a zero low half, and every step of the recurrence is one 16x2 add-max.  Lane l owns the CPL columns
// rbase + CPL*l + k.  CPL = 8: the general round (256 columns, rounds repeat until the window is covered); CPL = 12: windows
// of 256..383 columns in ONE round (a third of the rows of a 1-10 kb read; a second 8-column round costs as much as the first).
struct K2Packed { uint32_t c_mis, c_noe_del, c_noe_ins, c_ne_ins, c_eh, zero, mul4[4]; };

template <int GENERIC, int CPL, int VARIANT>
__device__ __forceinline__ void k2_round_packed(uint32_t* eh, const uint32_t* __restrict__ qs, uint32_t* zb, const int rm, const int nqw,
                                                const int rbase, const int j0, const int lim, const int fc, const int lane, const bool single,
                                                const uint32_t trep, const uint32_t rlo, const uint32_t rhi, const K2Packed& C, const int e_ins,
                                                int& carry, uint32_t& hcarry_pk, int& key, uint32_t& zlast)
{
    static_assert(CPL == 8 || (CPL == 12 && VARIANT == 1), "8 columns per lane, or 12 in single-round V1 rows");
    const int jl = rbase + CPL * lane;
    const int lo = j0 - jl, hi = lim - jl;                  // columns k with lo <= k < hi are cells of this row
    const int eC = CPL * e_ins;
    if (rbase < j0) {
        // first round of a window that does not start at a lane boundary: the columns left of j0 in lane 0 must not feed
        // the F chain.  They are dead for good (beg is monotone), so they are zeroed in the row buffer and their match
        // bits dropped: such a cell computes H = E = F = 0.  (Matrix-lookup scoring resets the chain per cell instead.)
        if (lane < j0 - rbase) eh[(rbase + lane) & rm] = 0u;
        __syncwarp();
    }
    uint32_t wd[CPL];
    {
        uint4 wa = make_uint4(0u, 0u, 0u, 0u), wb = wa, wc = wa;
        if (hi >= 0) {                                      // lanes right of the window read nothing (their columns may lie past the row buffer)
            wa = *reinterpret_cast<const uint4*>(eh + (jl & rm));
            wb = *reinterpret_cast<const uint4*>(eh + (jl & rm) + 4);
            if (CPL == 12) wc = *reinterpret_cast<const uint4*>(eh + ((jl + 8) & rm));      // a 12-column lane may straddle the end of the ring
            if (CPL == 12) wb = *reinterpret_cast<const uint4*>(eh + ((jl + 4) & rm));
        }
        wd[0] = wa.x; wd[1] = wa.y; wd[2] = wa.z; wd[3] = wa.w; wd[4] = wb.x; wd[5] = wb.y; wd[6] = wb.z; wd[7] = wb.w;
        if (CPL == 12) { wd[8] = wc.x; wd[9] = wc.y; wd[10] = wc.z; wd[11] = wc.w; }
    }
    // query nibbles of the lane's columns: xa = columns 0..7, xb = columns 8..11 (CPL 12: a lane starts at nibble 0 or 4 of a word)
    uint32_t xa, xb = 0;
    {
        const int qi = jl >> 3;
        const uint32_t q0 = qi < nqw ? qs[qi] : 0u;
        if (CPL == 8) xa = q0;
        else {
            const uint32_t q1 = qi + 1 < nqw ? qs[qi + 1] : 0u;
            const int sh = 4 * (jl & 7);
            xa = funnel_r(q0, q1, sh);
            xb = q1 >> sh;
        }
    }
    uint32_t ma0 = 0, ma1 = 0, mb = 0;                      // FAST: match bit of column k at bit 4*(k&3) of ma0 (k<4) / ma1 (k<8) / mb
    if (!GENERIC) {
        const uint32_t live = 0xffffffffu << (4 * imax(imin(lo, 8), 0));
        uint32_t y = xa ^ trep;
        y |= y >> 1; y |= y >> 2;
        ma0 = ~y & 0x11111111u & live;
        ma1 = ma0 >> 16;
        if (CPL == 12) {
            uint32_t z = xb ^ trep;
            z |= z >> 1; z |= z >> 2;
            mb = ~z & 0x1111u;                              // lo <= 7: columns 8..11 are never left of j0
        }
    }
    uint32_t hh[CPL], fl[CPL], t[CPL];
    uint32_t run = 0;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        uint32_t Wm;
        if (GENERIC) Wm = (uint32_t)k1_lookup(((k < 8 ? xa : xb) >> (4 * (k & 7))) & 15u, rlo, rhi) * 65536u + wd[k];
        else         Wm = ((k < 4 ? ma0 : (k < 8 ? ma1 : mb)) & (1u << (4 * (k & 3)))) * C.mul4[k & 3] + wd[k];
        uint32_t g;
        if (VARIANT == 2) {
            // upstream BWA: the zero guard "M ? M + s : 0" as one unsigned minimum (see BSW_K1_GUARD2 in bsw_k1_core.cuh),
            // both gap opens from M -- so t does not wait for F either
            const uint32_t mkp = add_max_s16x2(Wm, C.c_mis, C.zero);
            const uint32_t mk = umin32(mkp, 0u - (wd[k] & 0xffff0000u));
            hh[k] = max_s16x2(mk, Wm << 16);
            g = add_max_s16x2(mk, C.c_noe_ins, C.zero);
            t[k] = add_max_s16x2(mk, C.c_noe_del, C.zero);
        } else {
            hh[k] = add_max_s16x2(Wm, C.c_mis, Wm << 16);                          // {max(M + s, e), 0}  sx:1797,1798
            g = add_max_s16x2(hh[k], C.c_noe_ins, C.zero);                         // sx:1863,1865 with h >= hh
            if (GENERIC) { if (k == lo) run = 0; }                                 // (V2: a zeroed dead column yields M = 0 whatever the matrix says)
        }
        fl[k] = run;
        run = add_max_s16x2(run, C.c_ne_ins, g);                                   // sx:1780,1781
    }
    // carries across lanes: prefix max of run[l] + CPL*e_ins*l
    const int runi = (int)(run >> 16);
    int P = runi + eC * lane;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, P, d);
        if (lane >= d) P = imax(P, o);
    }
    const int Pex = __shfl_up_sync(0xffffffffu, P, 1);
    const int fin0 = (lane > 0) ? imax(Pex - eC * (lane - 1), 0) : 0;
    const int cin = imax(carry, 0);
    const int uin = imax(fin0, cin - eC * lane);                                   // f entering this lane's first column
    carry = __shfl_sync(0xffffffffu, imax(runi, uin - eC), 31);
    const uint32_t upk = (uint32_t)imax(uin, -1) << 16;                            // below zero it never wins: keep it inside 16 bits
    uint32_t hq[CPL];
    uint32_t nzacc = 0;
    int lkey = -1;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
        const uint32_t f = max_s16x2(fl[k], upk + (uint32_t)k * C.c_ne_ins);
        hq[k] = max_s16x2(hh[k], f);                                               // {h, 0}  sx:1809
        if (VARIANT == 1) t[k] = add_max_s16x2(hq[k], C.c_noe_del, C.zero);        // sx:1866,1862
        // sx:1808,1816: (h << 16) + k, ties to the right.  A zeroed column left of j0 yields h = 0 with +a/-b scoring and
        // under V2's zero guard, but max(s, 0) with a looked-up score under V1: there it is kept out of the arg-max
        // (it could win a row whose live cells are all below max(mat), e.g. the all-zero row that ends a task)
        if (!(GENERIC && VARIANT == 1) || k >= lo) lkey = add_max((int)hq[k], k, lkey);
        if (VARIANT == 1) nzacc += min(hq[k], 1u) << k;
    }
    // arg-max over ALL the lane's columns: a column right of the window can only win in the lane that holds `lim`; the
    // caller notices (column >= lim) and redoes the arg-max of this round from the row buffer
    if (hi > 0) key = imax(key, lkey + jl);
    uint32_t zbits = ~nzacc & ((1u << CPL) - 1u);                                  // V1: "h == 0" per column; cells outside the window are masked by the scan's ranges
    uint32_t hleft = __shfl_up_sync(0xffffffffu, hq[CPL - 1], 1);
    if (lane == 0) hleft = hcarry_pk;
    hcarry_pk = __shfl_sync(0xffffffffu, hq[CPL - 1], 31);
    if (hi >= 0 && lo < CPL) {
        uint32_t ow[CPL];
        ow[0] = add_max_s16x2(wd[0], C.c_eh, pack_hi_hi(hleft, t[0]));             // {H(i, j-1), max(e - e_del, t)}  sx:1770-1771,1776
#pragma unroll
        for (int k = 1; k < CPL; ++k) ow[k] = add_max_s16x2(wd[k], C.c_eh, pack_hi_hi(hq[k - 1], t[k]));
        if (VARIANT == 2 && !(lo < 0 && hi >= CPL)) {
            // V2 boundary lane: exactly the cells [lo, hi) and the end slot, nothing beside them -- the zero scan may grow the
            // window, and a later row then reads the slots next to it as they are.  First cell: H half = fc; end slot: E half = 0.
            uint32_t nz = 0;
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                if (k >= lo && k <= hi) {
                    uint32_t v = (k == lo) ? ((ow[k] & 0x0000ffffu) | ((uint32_t)fc << 16)) : ow[k];
                    if (k == hi) v &= 0xffff0000u;
                    eh[(jl + k) & rm] = v;
                    nz |= (v != 0u ? 1u : 0u) << k;
                }
            }
            zbits = nz;
        } else {
            // every lane that touches [j0, lim] stores its words; a V1 boundary lane patches two of them afterwards: the
            // first cell's left neighbour is the first-column value fc, the end slot eh[lim] is {h1, 0} (sx:1775,1904,1776);
            // what lands left of j0 or right of lim is never read
            *reinterpret_cast<uint4*>(eh + (jl & rm)) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            *reinterpret_cast<uint4*>(eh + ((jl + 4) & rm)) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
            if (CPL == 12) *reinterpret_cast<uint4*>(eh + ((jl + 8) & rm)) = make_uint4(ow[8], ow[9], ow[10], ow[11]);
            if (VARIANT == 1) {
                unsigned short* eh16 = reinterpret_cast<unsigned short*>(eh);
                if (lo > 0) eh16[2 * (j0 & rm) + 1] = (unsigned short)fc;          // H half of column j0 (lo == 0: hleft already is fc)
                if (hi < CPL) eh16[2 * (lim & rm)] = 0;                            // E half of the end slot
            } else {
                uint32_t nz = 0;                                                   // V2 keeps "word != 0" per column (full lane: all are cells)
#pragma unroll
                for (int k = 0; k < CPL; ++k) nz += min(ow[k], 1u) << k;
                zbits = nz;
            }
        }
        if (CPL == 8 && !single) reinterpret_cast<unsigned char*>(zb)[(jl & rm) >> 3] = (unsigned char)zbits;
    }
    zlast = zbits;
}

template <int NW> __device__ __forceinline__ void k2_sync() { if (NW == 1) __syncwarp(); else __syncthreads(); }

// VARIANT 1 = the RTL's recurrence, 2 = upstream BWA's (SURVEY appendix B "V2 deltas"): the zero guard on M, gap opens
// taken from M, the first-column value only while beg == 0, and the zero-scan narrowing.  V2's narrowing can grow the
// window past the cells the previous row wrote (end' = j + 2), so the next row reads a STALE slot of the row buffer --
// whatever an older row, or the first-row fill, left there.  Two consequences here: boundary lanes store exactly the
// cells [j0, lim] (V1 may scribble over the columns next to the window, nobody reads them), and in ring mode a column
// that was never written is given its first-row value before the first row that can read it.
template <int GENERIC, int K2_WARPS, int VARIANT>
__global__ void __launch_bounds__(32 * K2_WARPS, K2_WARPS == 1 ? 24 : 1) k2_extend_kernel(const __grid_constant__ LaunchArgs A)
{
    constexpr int K2_NT = 32 * K2_WARPS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TileHdr hd = A.tiles[blockIdx.x];
    const uint32_t slot = hd.slot0;
    const SlotParam sp = A.slots[slot];
    const int qlen = sp.qlen, tlen = sp.tlen, h0 = sp.h0, w = sp.w;
    const int qcap = (A.qmax + 1 + K2_GROUP - 1) & ~(K2_GROUP - 1);      // query columns, multiple of 256
    // The row buffer only has to hold the live window [beg & ~255, end] of a row (at most 2w+1+255 columns), so for long
    // queries it is a ring of K2_RING columns indexed by (column & rm); groups are 256-aligned, so a group never wraps.
    const bool ring = qcap > K2_RING && 2 * A.wmax + 1 + 2 * K2_GROUP <= K2_RING;
    const int rcap = ring ? K2_RING : qcap;
    const int rm = ring ? (K2_RING - 1) : 0x7fffffff;
    const int nqw = (qlen + 7) >> 3;
    const uint32_t qbytes = (uint32_t)((nqw * 4 + 15) & ~15);

    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    int* xagg = reinterpret_cast<int*>(smem_raw + 16);                            // [K2_WARPS] group aggregate of pass 1
    int* xhl  = xagg + K2_WARPS;                                                  // [K2_WARPS] h of the group's last column
    int* xkey = xhl + K2_WARPS;                                                   // [K2_WARPS] per-warp arg-max key
    // Short queries are staged in shared memory by TMA; in ring mode (long queries) the packed query stays in the source
    // arena: lane l of a group reads word (column >> 3), consecutive lanes read consecutive words, L1/L2 resident.
    uint32_t* qsm = reinterpret_cast<uint32_t*>(smem_raw + K2_HDR_BYTES);         // qcap/8 words (+ pad), unused in ring mode
    const uint32_t* qs = ring ? (A.arena + (size_t)hd.qoff16 * 4u) : qsm;
    uint32_t* zb = qsm + (ring ? 0 : (qcap >> 3)) + 4;                            // rcap/32 words of zero bits
    uint32_t* eh = zb + (rcap >> 5) + 4;                                          // rcap + 16 words
    eh = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(eh) + 15) & ~(uintptr_t)15);

    if (tid == 0 && !ring) {
        const uint32_t bar = k2_smem_u32(mbar), dst = k2_smem_u32(qsm);
        const void* src = reinterpret_cast<const uint4*>(A.arena) + hd.qoff16;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(qbytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(qbytes), "r"(bar) : "memory");
    }

    const int o_del = A.p.o_del, e_del = A.p.e_del, e_ins = A.p.e_ins;
    const int oe_del = A.p.o_del + A.p.e_del, oe_ins = A.p.o_ins + A.p.e_ins;
    const int zdrop = A.p.zdrop;
    const int mat = A.p.match, mis = -A.p.mismatch;
    // One warp per task, V1: the packed cell of K1 (bsw_k1_core.cuh) -- the row-buffer word is {H hi16, E lo16}, H / F / the
    // running maxima stay in the high half of a register with a zero low half, and every step of the recurrence is one
    // 16x2 add-max.  The other instantiations keep {E hi16, H lo16} and the scalar cell.
    constexpr bool HI = (K2_WARPS == 1);
    const uint32_t ce_pack = HI ? (0x80000000u | ((uint32_t)(-e_del) & 0xffffu)) : (0x8000u | ((uint32_t)(-e_del) << 16));
    const int e8 = 8 * e_ins, e256 = K2_GROUP * e_ins;
    K2Packed PK;
    PK.c_mis = ((uint32_t)(GENERIC ? 0 : mis) << 16) | 0x8000u;                  // {-b (or 0), -32768}
    PK.c_noe_del = (uint32_t)(-oe_del) << 16;                                    // {-oe_del, 0}
    PK.c_noe_ins = (uint32_t)(-oe_ins) << 16;                                    // {-oe_ins, 0}
    PK.c_ne_ins = (uint32_t)(-e_ins) << 16;                                      // {-e_ins, 0}
    PK.c_eh = 0x80000000u | ((uint32_t)(-e_del) & 0xffffu);                      // {-32768, -e_del}
    PK.zero = A.p.zero;                                                          // 0, opaque to ptxas
#pragma unroll
    for (int k = 0; k < 4; ++k) PK.mul4[k] = (uint32_t)(mat - mis) << (16 - 4 * k);   // match bit 4k -> +(a+b) in the H half
    if (HI) {
        asm volatile("" : "+r"(PK.c_mis), "+r"(PK.c_noe_del), "+r"(PK.c_noe_ins), "+r"(PK.c_ne_ins), "+r"(PK.c_eh), "+r"(PK.zero));
        asm volatile("" : "+r"(PK.mul4[0]), "+r"(PK.mul4[1]), "+r"(PK.mul4[2]), "+r"(PK.mul4[3]));
    }

    // first row: eh[j].h = H(-1, j-1), all e = 0 (sx:1818; 1979,1957,1974; 1975-1978,1821)
    for (int j = tid; j < rcap + 16; j += K2_NT) {
        int hv = (j == 0) ? h0 : imax(h0 - A.p.o_ins - j * e_ins, 0);
        if (j > qlen) hv = 0;
        eh[j] = HI ? ((uint32_t)hv << 16) : (uint32_t)hv;
    }
    __syncthreads();                                   // mbarrier init + first row visible to every warp
    if (!ring) {
        const uint32_t bar = k2_smem_u32(mbar);
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar) : "memory");
        }
    }

    const uint32_t* tg = A.arena + (size_t)hd.toff16 * 4u;
    int max = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0;   // sx:889,1009,919,1019,1029,929
    int beg = 0, end = qlen;                                                       // sx:769,779
    unsigned long long cells = 0;
    uint32_t tw = 0;
    int hiw = rcap - 1;                  // V2: highest column whose slot holds its own (written or first-row) value

    for (int i = 0; i < tlen; ++i) {                                               // sx:1891
        if ((i & 7) == 0) tw = __ldg(tg + (i >> 3));                               // warp-uniform broadcast load
        const uint32_t tb = (tw >> ((i & 7) * 4)) & 15u;
        const uint32_t trep = tb * 0x11111111u;
        uint32_t rlo = 0, rhi = 0;
        if (GENERIC) { rlo = A.p.row_lo[tb]; rhi = A.p.row_hi[tb]; }

        const int j0 = imax(beg, i - w);                                           // sx:1846,1894,1895,1803
        const int lim = imin(imin(end, i + w + 1), qlen);                          // sx:1980,1843,1897,1898,1842
        const int fc = (VARIANT == 1 || j0 == 0) ? imax(h0 - (o_del + e_del * (i + 1)), 0) : 0;    // V1: unconditional (sx:1796,1795,1880,1835,849)
        if (VARIANT == 2 && lim - 1 > hiw) {
            // ring mode: columns the window reaches for the first time still hold what column - 2048 left there; a
            // flat row buffer would show the first-row fill (sx:1975-1978) -- put it there
            for (int c = hiw + 1 + tid; c <= lim - 1; c += K2_NT) eh[c & rm] = (uint32_t)(c > qlen ? 0 : imax(h0 - A.p.o_ins - c * e_ins, 0)) << (HI ? 16 : 0);
            k2_sync<K2_WARPS>();
        }
        if (VARIANT == 2) hiw = imax(hiw, lim);
        if (lim <= j0) {
            // empty row: the reference's loop body never runs, h1 = fc, j stays at beg, m == 0 -> break
            if (j0 == qlen) { if (!(gscore > fc)) { max_ie = i; gscore = fc; } }   // sx:1768,1913,1941
            break;                                                                 // sx:1942
        }

        int key = -1, h1 = 0, cb = -1, ce = 0x7fffffff;
        const bool narrow = VARIANT == 1 && K2_WARPS == 1 && A.k2_narrow && lim - j0 < 64;           // warp-uniform
        if (narrow) {
            if (lim - j0 < 32)
                k2_narrow_row<GENERIC, 1, HI>(eh, qs, rm, nqw, j0, lim, fc, lane, trep, rlo, rhi, mat, mis, e_ins, oe_ins, oe_del, ce_pack, key, h1, cb, ce);
            else
                k2_narrow_row<GENERIC, 2, HI>(eh, qs, rm, nqw, j0, lim, fc, lane, trep, rlo, rhi, mat, mis, e_ins, oe_ins, oe_del, ce_pack, key, h1, cb, ce);
            __syncwarp();                                                          // the row buffer is read by other lanes in the next row
        } else {
        int carry = 0;            // f entering the first column of the round
        int hcarry = fc;          // h of the column left of the round
        // one warp, one round (the window fits 256 columns from j0 & ~7 -- nearly every row): the zero bits stay in
        // registers and the narrowing below needs no pass over shared memory
        const int span = lim - (j0 & ~7);
        const bool wide12 = HI && VARIANT == 1 && span >= K2_GROUP && span < 12 * 32;          // warp-uniform: one round of 12 columns per lane
        const bool single = VARIANT == 1 && K2_WARPS == 1 && (span < K2_GROUP || wide12);
        uint32_t zlast = 0;
        uint32_t hcarry_pk = (uint32_t)fc << 16;
        int keyprev = -1, jl_last = 0, cpl_last = 8;
        if constexpr (HI) {
            if (wide12) {
                jl_last = (j0 & ~7) + 12 * lane; cpl_last = 12;
                k2_round_packed<GENERIC, 12, 1>(eh, qs, zb, rm, nqw, j0 & ~7, j0, lim, fc, lane, true, trep, rlo, rhi, PK, e_ins, carry, hcarry_pk, key, zlast);
            }
        }
        if (!wide12)
        // rounds start at the window (rounded down to a lane's 8 columns), not at a 256-column boundary: the live window
        // of a 1-10 kb PacBio-like task is ~200 columns wide (tools: 3.2 M rows, mean 203, 99 % below 384), and an aligned
        // group would split about half of those rows into two rounds
        for (int rbase = j0 & ~7; rbase <= lim; rbase += K2_GROUP * K2_WARPS) {
            const int gbase = rbase + K2_GROUP * warp;
            const bool active = gbase <= lim;                   // warp-uniform
            const int jl = gbase + 8 * lane;
            const int lo = j0 - jl, hi = lim - jl;              // columns k with lo <= k < hi are cells of this row
            const bool full = (lo < 0) && (hi >= 8);            // lo == 0: the lane holds the first live cell (its left neighbour is fc)
            const int klo = imax(lo, 0), khi = imin(hi, 8);
            const uint32_t livebits = khi > klo ? ((0xffu >> (8 - khi)) & (0xffu << klo)) : 0u;     // bit k: column jl + k is a cell of this row
            if constexpr (HI) {
                keyprev = key; jl_last = jl; cpl_last = 8;
                k2_round_packed<GENERIC, 8, VARIANT>(eh, qs, zb, rm, nqw, rbase, j0, lim, fc, lane, single, trep, rlo, rhi, PK, e_ins, carry, hcarry_pk, key, zlast);
            } else {
            uint32_t wd[8];
            int hh[8], fl[8], mk[VARIANT == 2 ? 8 : 1];
            int run = 0, fin0 = 0;
            if (active) {
                uint4 wa = make_uint4(0u, 0u, 0u, 0u), wb = wa;
                if (hi >= 0) {                                  // lanes right of the window read nothing (rounds are not 256-aligned:
                    wa = *reinterpret_cast<const uint4*>(eh + (jl & rm));          // their columns may lie past the row buffer)
                    wb = *reinterpret_cast<const uint4*>(eh + (jl & rm) + 4);
                }
                wd[0] = wa.x; wd[1] = wa.y; wd[2] = wa.z; wd[3] = wa.w; wd[4] = wb.x; wd[5] = wb.y; wd[6] = wb.z; wd[7] = wb.w;
                const uint32_t qw = (jl >> 3) < nqw ? qs[jl >> 3] : 0u;
                const uint32_t x = GENERIC ? qw : (qw ^ trep);
                // pass 1: everything that does not need the incoming F
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int M = (int)(wd[k] & 0xffffu), e = (int)(wd[k] >> 16);
                    const int s = k2_score<GENERIC>(GENERIC ? ((x >> (4 * k)) & 15u) : (x & (0xfu << (4 * k))), mat, mis, rlo, rhi);
                    int g;
                    if (VARIANT == 2) {
                        mk[k] = M ? M + s : 0;                                         // V2: zero guard
                        hh[k] = imax(mk[k], e);
                        g = add_max_relu(mk[k], -oe_ins, 0);                           // V2: gap open from M
                    } else {
                        hh[k] = add_max(M, s, e);                                      // sx:1797,1798
                        g = add_max_relu(hh[k], -oe_ins, 0);                           // sx:1863,1865 with h >= hh
                    }
                    // Cells outside [j0, lim) are computed like the others and simply never looked at: columns left of j0
                    // are dead (beg is monotone), columns right of lim are rewritten before they are read (every row
                    // ends by writing its end slot), and F only flows to the right -- so the one thing to protect is the
                    // F chain of the first live cell, which must start from zero.
                    if (k == lo) run = 0;
                    fl[k] = run;
                    run = add_max(run, -e_ins, g);                                     // sx:1780,1781
                }
                // carries across lanes: prefix max of A[l] + 8*e_ins*l
                int P = run + e8 * lane;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int o = __shfl_up_sync(0xffffffffu, P, d);
                    if (lane >= d) P = imax(P, o);
                }
                const int Pex = __shfl_up_sync(0xffffffffu, P, 1);
                fin0 = (lane > 0) ? imax(Pex - e8 * (lane - 1), 0) : 0;    // f entering this lane if the group's carry-in were 0
                if (K2_WARPS > 1 && lane == 31) xagg[warp] = imax(run, fin0 - e8);     // f leaving the group (zero carry-in)
            } else if (K2_WARPS > 1 && lane == 31) {
                xagg[warp] = 0;
            }
            int cin;
            if (K2_WARPS == 1) {
                // one warp per task: the round's carry stays in registers (a group is always active here)
                cin = imax(carry, 0);
                const int fout = imax(run, imax(fin0, cin - e8 * lane) - e8);
                carry = __shfl_sync(0xffffffffu, fout, 31);
            } else {
                k2_sync<K2_WARPS>();                                       // (1) group aggregates visible
                // f entering this warp's group: the round's carry and the aggregates of the groups before it
                cin = carry - e256 * warp;
                int cnext = carry - e256 * K2_WARPS;
#pragma unroll
                for (int g = 0; g < K2_WARPS; ++g) {
                    const int ag = xagg[g];
                    if (g < warp) cin = imax(cin, ag - e256 * (warp - 1 - g));
                    cnext = imax(cnext, ag - e256 * (K2_WARPS - 1 - g));
                }
                cin = imax(cin, 0);
                carry = imax(cnext, 0);
            }

            int h[8];
            uint32_t enew[8];
            uint32_t zbits = 0;
            if (active) {
                // pass 2: fold the carry, finish H, E, key, zero bits
                int u = imax(fin0, cin - e8 * lane);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int f = imax(fl[k], u);
                    u -= e_ins;
                    h[k] = imax(hh[k], f);                                             // sx:1809
                    const int t = add_max_relu(VARIANT == 2 ? mk[k] : h[k], -oe_del, 0);   // sx:1866,1862
                    enew[k] = add_max_s16x2(wd[k], ce_pack, (uint32_t)t << 16);        // {max(e-e_del,t), max(M-32768,0)=0}  sx:1770-1771
                    const int kk = h[k] * 65536 + jl + k;                              // sx:1808,1816
                    key = imax(key, (livebits & (1u << k)) ? kk : -1);
                    zbits |= (uint32_t)(1 - imin(h[k], 1)) << k;                       // cells outside the window are masked by the scan's ranges
                }
                if (K2_WARPS > 1 && lane == 31) xhl[warp] = h[7];
            }
            if (K2_WARPS > 1) k2_sync<K2_WARPS>();                         // (2) last-column h of every group visible
            if (active) {
                int hleft = __shfl_up_sync(0xffffffffu, h[7], 1);
                if (lane == 0) hleft = (K2_WARPS == 1 || warp == 0) ? hcarry : xhl[warp - 1];
                if (K2_WARPS == 1) hcarry = __shfl_sync(0xffffffffu, h[7], 31);
                if (hi >= 0 && lo < 8) {
                    // every lane that touches [j0, lim] stores its 8 words; a boundary lane patches two of them first:
                    // the first cell's left neighbour is the first-column value fc, the end slot eh[lim] is {h1, 0}
                    // (sx:1775,1904,1776); what lands left of j0 or right of lim is never read (see pass 1)
                    uint32_t ow[8];
                    ow[0] = enew[0] | (uint32_t)hleft;
#pragma unroll
                    for (int k = 1; k < 8; ++k) ow[k] = enew[k] | (uint32_t)h[k - 1];
                    if (VARIANT == 2 && !full) {
                        // V2 boundary lane: exactly the cells [lo, hi) and the end slot, nothing beside them (stale slots are read later)
                        uint32_t nz = 0;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            if (k >= lo && k <= hi) {
                                uint32_t v = (k == lo) ? ((ow[k] & 0xffff0000u) | (uint32_t)fc) : ow[k];
                                if (k == hi) v &= 0x0000ffffu;
                                eh[(jl + k) & rm] = v;
                                nz |= (v != 0u ? 1u : 0u) << k;
                            }
                        }
                        zbits = nz;
                    } else {
                    *reinterpret_cast<uint4*>(eh + (jl & rm)) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                    *reinterpret_cast<uint4*>(eh + (jl & rm) + 4) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
                    // the two patches are 16-bit stores behind the vector stores (same thread: ordered)
                    unsigned short* eh16 = reinterpret_cast<unsigned short*>(eh);
                    if (lo > 0) eh16[2 * (j0 & rm)] = (unsigned short)fc;           // H half of column j0 (lo == 0: hleft already is fc)
                    if (hi < 8) eh16[2 * (lim & rm) + 1] = 0;                       // E half of the end slot
                    if (VARIANT == 2) {                                             // V2 keeps "word != 0" bits per column (full lane: all eight are cells)
                        zbits = 0;
#pragma unroll
                        for (int k = 0; k < 8; ++k) zbits |= (ow[k] != 0u ? 1u : 0u) << k;
                    }
                    }
                    if (!single) reinterpret_cast<unsigned char*>(zb)[(jl & rm) >> 3] = (unsigned char)zbits;
                }
                zlast = zbits;
            }
            if (K2_WARPS > 1) hcarry = xhl[K2_WARPS - 1];      // only consumed when another round follows (then the last warp was active)
            }   // scalar cell
        }
        key = __reduce_max_sync(0xffffffffu, key);
        if (K2_WARPS > 1 && lane == 0) xkey[warp] = key;
        k2_sync<K2_WARPS>();                                               // (3) row buffer, zero bits and keys visible
        if (HI && (key & 0xffff) >= lim) {
            // the packed round takes the arg-max over all the columns of a lane; the lane that holds `lim` also holds columns
            // right of the window, and one of them won: redo the last round's share with those columns left out, from the
            // row buffer (slot c + 1 holds h of column c).  Taken by 5-15 % of the rows (stale H values right of a window
            // that has just moved left are often larger than anything in it); ~40 instructions when it is.
            int mk2 = keyprev;
            for (int k = 0; k < cpl_last; ++k) {
                const int c = jl_last + k;
                if (c >= j0 && c < lim) mk2 = imax(mk2, (int)(eh[(c + 1) & rm] & 0xffff0000u) + c);
            }
            key = __reduce_max_sync(0xffffffffu, mk2);
        }
        if (K2_WARPS > 1) {
            key = xkey[0];
#pragma unroll
            for (int g = 1; g < K2_WARPS; ++g) key = imax(key, xkey[g]);
        }
        h1 = HI ? (int)(eh[lim & rm] >> 16) : (int)(eh[lim & rm] & 0xffffu);
        if (single) {
            // narrowing from the lanes' own zero bits (sx:1766-1769 / 1779,1782-1789): bit k of zlast <-> h of column jl + k is 0
            const int mjw = key & 0xffff, jl = (j0 & ~7) + cpl_last * lane;
            const uint32_t za = zlast & k2_range_mask(jl, j0, mjw - 1);
            const uint32_t ze = zlast & k2_range_mask(jl, mjw + 1, lim - 1);
            cb = __reduce_max_sync(0xffffffffu, za ? jl + 31 - __clz(za) : -1);
            ce = __reduce_min_sync(0xffffffffu, ze ? jl + __ffs(ze) - 1 : 0x7fffffff);
        } else if (VARIANT == 2) {
            // BWA's zero scan over the stored words of columns [j0, lim]: cb = first non-zero column in [j0, lim-1],
            // ce = last non-zero column in [beg', lim]; bit k of a lane's byte <-> the word of column jl + k is non-zero
            int first = 0x7fffffff;
            for (int wbase = (j0 >> 5); wbase <= (lim >> 5); wbase += 32) {
                const int wi = wbase + lane;
                const uint32_t zw = (wi <= (lim >> 5)) ? zb[wi & (rm >> 5)] : 0u;
                const uint32_t za = zw & k2_range_mask(wi * 32, j0, lim - 1);
                if (za) first = imin(first, wi * 32 + __ffs(za) - 1);
            }
            first = __reduce_min_sync(0xffffffffu, first);
            const int nbeg = first != 0x7fffffff ? first : lim;
            int last = -1;
            for (int wbase = (nbeg >> 5); wbase <= (lim >> 5); wbase += 32) {
                const int wi = wbase + lane;
                const uint32_t zw = (wi <= (lim >> 5)) ? zb[wi & (rm >> 5)] : 0u;
                const uint32_t ze = zw & k2_range_mask(wi * 32, nbeg, lim);
                if (ze) last = imax(last, wi * 32 + 31 - __clz(ze));
            }
            last = __reduce_max_sync(0xffffffffu, last);
            cb = nbeg;                                                             // carried to the common code below
            ce = imin((last >= 0 ? last : nbeg - 1) + 2, qlen);
        } else {
            // narrowing scan (sx:1766-1769 / 1779,1782-1789): zero bit of cell c <-> eh[c+1].h == 0
            const int mjw = key & 0xffff;
            for (int wbase = (j0 >> 5); wbase <= ((lim - 1) >> 5); wbase += 32) {
                const int wi = wbase + lane;
                const uint32_t zw = (wi <= ((lim - 1) >> 5)) ? zb[wi & (rm >> 5)] : 0u;
                const uint32_t za = zw & k2_range_mask(wi * 32, j0, mjw - 1);
                const uint32_t ze = zw & k2_range_mask(wi * 32, mjw + 1, lim - 1);
                if (za) cb = imax(cb, wi * 32 + 31 - __clz(za));
                if (ze) ce = imin(ce, wi * 32 + __ffs(ze) - 1);
            }
            cb = __reduce_max_sync(0xffffffffu, cb);
            ce = __reduce_min_sync(0xffffffffu, ce);
        }
        }   // general path

        // ---- row epilogue (identical in every thread of the CTA) ----
        const int m = key >> 16, mj = key & 0xffff;
        cells += (unsigned long long)(lim - j0);
        if (lim == qlen) {                                                         // sx:1768,1913
            if (!(gscore > h1)) { max_ie = i; gscore = h1; }                       // sx:1941,1829,1831
        }
        if (m == 0) break;                                                         // sx:1942
        if (m > max) {                                                             // sx:1959
            max = m; max_i = i; max_j = mj;
            const int d = mj > i ? mj - i : i - mj;
            max_off = max_off > d ? max_off : d;                                   // sx:1707-1708,1812
        } else if (zdrop > 0) {                                                    // ksw_extend2 z-drop (not in the RTL)
            const int di = i - max_i, dj = mj - max_j;
            if (di > dj) { if (max - m - (di - dj) * e_del > zdrop) break; }
            else         { if (max - m - (dj - di) * e_ins > zdrop) break; }
        }
        // narrowing: beg' = 2 + last zero cell in [j0, mj-1], else (fc == 0 ? j0+1 : j0)
        //            end' = 1 + first zero cell in [mj+1, lim-1], else lim+1
        (void)mj;
        if (VARIANT == 2) { beg = cb; end = ce; continue; }
        beg = cb >= 0 ? cb + 2 : (fc == 0 ? j0 + 1 : j0);
        end = ce != 0x7fffffff ? ce + 1 : lim + 1;
    }

    if (tid == 0) {
        int4* o = reinterpret_cast<int4*>(A.out + (A.out_index ? A.out_index[slot] : slot));
        const unsigned long long cc = cells > 0x7fffffffull ? 0x7fffffffull : cells;
        o[0] = make_int4(max, max_j + 1, max_i + 1, max_ie + 1);                   // sx:1315-1375 (score,qle,tle,gtle)
        o[1] = make_int4(gscore, max_off, (int)cc, STATUS_OK);
        if (A.cells_total && cells) atomicAdd(A.cells_total, cells);
    }
}

size_t k2_smem_bytes(int qmax, int wmax)
{
    const size_t qcap = ((size_t)qmax + 1 + K2_GROUP - 1) & ~(size_t)(K2_GROUP - 1);
    const bool ring = qcap > (size_t)K2_RING && 2 * (size_t)wmax + 1 + 2 * K2_GROUP <= (size_t)K2_RING;
    const size_t rcap = ring ? (size_t)K2_RING : qcap;
    return (size_t)K2_HDR_BYTES + ((ring ? 0 : (qcap >> 3)) + 4 + (rcap >> 5) + 4 + rcap + 16) * 4u + 16u;
}

template <int GENERIC, int NW, int VARIANT>
static cudaError_t k2_launch_t(const LaunchArgs& a, cudaStream_t st)
{
    if (!a.ntiles) return cudaSuccess;
    const size_t smem = k2_smem_bytes(a.qmax, a.wmax);
    auto kern = k2_extend_kernel<GENERIC, NW, VARIANT>;
    // always the same (maximal) value: launches are issued concurrently from several host threads, and a per-launch
    // value would race with another thread's launch of the same kernel
    if (smem > 232448) return cudaErrorInvalidValue;
    static std::atomic<unsigned> smem_set{ 0u };          // per instantiation of this launcher
    cudaError_t err = ensure_max_smem(kern, smem_set);
    if (err != cudaSuccess) return err;
    kern<<<a.ntiles, 32 * NW, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t k2_launch(const LaunchArgs& a, int generic, int warps, int variant, cudaStream_t st)
{
    if (variant == 2) return generic ? k2_launch_t<1, 1, 2>(a, st) : k2_launch_t<0, 1, 2>(a, st);      // V2: one warp per task
    if (warps >= K2_MAXW) return generic ? k2_launch_t<1, 4, 1>(a, st) : k2_launch_t<0, 4, 1>(a, st);
    return generic ? k2_launch_t<1, 1, 1>(a, st) : k2_launch_t<0, 1, 1>(a, st);
}

}  // namespace bsw
