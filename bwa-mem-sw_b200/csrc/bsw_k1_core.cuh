// bsw_k1_core.cuh -- one extension task as executed by one lane of the inter-task kernel K1.
//
// This is the body of what one FPGA PE runs for one sw_extend call (sw_pe_array_sw_extend.v FSM
// :1639-1705).  It is written as a __host__ __device__ function over the K1 tile layout
// (bsw_device.cuh) so that csrc/emu.cpp can execute the *identical* control flow and arithmetic on the
// CPU (tests only; the product library contains only the device instantiation).
//
// Row state.  BWA's eh_t eh[qlen+1] / the RTL's eh_arr ({E[15:8],H[7:0]}, sw_pe_array_sw_extend_eh_arr.v)
// is one 32-bit word {H[31:16], E[15:0]} per column in shared memory at eh[j*32 + lane]: a warp's 32
// tasks always hit 32 different banks, whatever column each lane is at.  Scores are < 32768 by the
// scheduler's admission rule (h0 + qlen*max(mat) <= 32767), so 16 bits are exact.
//
// Branch-free cell (8-column chunks whose H are all non-zero -- the common case inside the band).  The whole
// cell stays in packed 16x2 form, H in the high half, so nothing is ever unpacked:
//     v   = eq & (1<<k)                              LOP3    match bit of this column (one-hot query planes)
//     Wm  = v * ((a+b) << (16-k)) + W                IMAD    {M + (a+b)*match, e}             (FMA pipe)
//     C   = Wm << 16                                 IMAD.SHL {e, 0}                          (FMA pipe)
//     hh  = max(Wm + {-b, -32768}, C)                VIADDMNMX.S16x2  {max(M+s, e), 0}        sx:1797,1798
//     h   = max(hh, f)                               VIMNMX.S16x2     {h, 0}                   sx:1809
//     t   = max(h + {-oe_del, 0}, 0)                 VIADDMNMX.S16x2.RELU {t, 0}               sx:1866,1862
//     eh  = max(W + {-32768, -e_del}, {h1, t})       PRMT + VIADDMNMX.S16x2 -> {h1, max(e-e_del,t)}  sx:1776,1770-1771
//     f   = max(f + {-e_ins, 0}, t)                  VIADDMNMX.S16x2                           sx:1863,1865,1780-1781
//     key = max(key, h + k)                          VIADDMNMX  (h is already h<<16: row max + right-most column) sx:1808,1816
// = 8 ALU-pipe + 2 FMA-pipe instructions, 1 LDS, 1 STS per cell for 13 algorithmic integer ops.
// The loop-carried chain of a row is f -> h -> t -> f (three dependent instructions per cell).  Opening the gap from hh
// instead (max(f - e_ins, h - oe_ins) = max(f - e_ins, hh - oe_ins), since f - oe_ins <= f - e_ins) cuts it to one, but
// with symmetric gap penalties it costs a ninth ALU instruction (t no longer serves both gaps): measured 1 110 against
// 1 135-1 142 GCUPS on 1 M x 150 bp (r02) -- the kernel is bound by ALU issue, not by the chain.  The asymmetric path
// needs its own instruction anyway and takes it from hh.
//
// Band narrowing (V1).  The reference recomputes [beg,end) after every row by scanning the stored
// row for the run of non-zero H around mj (sw_pe_array_sw_extend.v:1766-1769,1779,1782-1789).  A
// second pass over the row would double the shared-memory traffic, so K1 applies the same rule lazily
// inside the NEXT row, which reads every eh[j].h of the candidate window [beg, end+1) anyway:
//   * a zero at column j <= mj means beg' >= j+1  -> the row restarts at j+1 (f=0, h1=first column)
//   * a zero at column j >= mj+2 means end' = j    -> the row ends at j
// Cells evaluated before a restart only touch eh[] slots left of beg', which are never read again
// (beg is monotone), so the visible result -- including the cells count -- is bit-identical.
#pragma once
#include "bsw_device.cuh"

#if defined(__CUDACC__)
#define BSW_HD __host__ __device__ __forceinline__
#else
#define BSW_HD inline
#endif

namespace bsw {

// ---- integer helpers: DPX intrinsics on the device, plain C on the host emulation ----
BSW_HD int imax(int a, int b) { return a > b ? a : b; }
BSW_HD int imin(int a, int b) { return a < b ? a : b; }

BSW_HD int add_max(int a, int b, int c)        // max(a + b, c)
{
#if defined(__CUDA_ARCH__)
    return __viaddmax_s32(a, b, c);
#else
    return imax(a + b, c);
#endif
}
BSW_HD int add_max_relu(int a, int b, int c)   // max(a + b, c, 0)
{
#if defined(__CUDA_ARCH__)
    return __viaddmax_s32_relu(a, b, c);
#else
    return imax(imax(a + b, c), 0);
#endif
}
#if !defined(__CUDA_ARCH__)
inline int16_t bsw_lo16(uint32_t x) { return (int16_t)(uint16_t)(x & 0xffffu); }
inline int16_t bsw_hi16(uint32_t x) { return (int16_t)(uint16_t)(x >> 16); }
inline uint32_t bsw_mk16x2(int16_t hi, int16_t lo) { return ((uint32_t)(uint16_t)hi << 16) | (uint32_t)(uint16_t)lo; }
#endif
BSW_HD uint32_t add_max_s16x2(uint32_t a, uint32_t b, uint32_t c)   // per 16-bit half: max(a + b, c), wrapping add
{
#if defined(__CUDA_ARCH__)
    return __viaddmax_s16x2(a, b, c);
#else
    const int16_t lo = (int16_t)(uint16_t)((a & 0xffffu) + (b & 0xffffu)), hi = (int16_t)(uint16_t)((a >> 16) + (b >> 16));
    return bsw_mk16x2(hi > bsw_hi16(c) ? hi : bsw_hi16(c), lo > bsw_lo16(c) ? lo : bsw_lo16(c));
#endif
}
BSW_HD uint32_t add_max_s16x2_relu(uint32_t a, uint32_t b, uint32_t c)   // per half: max(a + b, c, 0)
{
#if defined(__CUDA_ARCH__)
    return __viaddmax_s16x2_relu(a, b, c);
#else
    const uint32_t r = add_max_s16x2(a, b, c);
    return bsw_mk16x2(bsw_hi16(r) > 0 ? bsw_hi16(r) : 0, bsw_lo16(r) > 0 ? bsw_lo16(r) : 0);
#endif
}
BSW_HD uint32_t max_s16x2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __vmaxs2(a, b);
#else
    return bsw_mk16x2(bsw_hi16(a) > bsw_hi16(b) ? bsw_hi16(a) : bsw_hi16(b), bsw_lo16(a) > bsw_lo16(b) ? bsw_lo16(a) : bsw_lo16(b));
#endif
}
BSW_HD uint32_t min3_u16x2(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    return __vimin3_u16x2(a, b, c);
#else
    uint32_t lo = a & 0xffffu, hi = a >> 16;
    if ((b & 0xffffu) < lo) lo = b & 0xffffu;
    if ((c & 0xffffu) < lo) lo = c & 0xffffu;
    if ((b >> 16) < hi) hi = b >> 16;
    if ((c >> 16) < hi) hi = c >> 16;
    return lo | (hi << 16);
#endif
}
BSW_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, int sh)   // ({hi,lo} >> sh) low word, 0 <= sh < 32
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}
BSW_HD uint32_t pack_hi_hi(uint32_t hi_src, uint32_t lo_src)   // {hi_src[31:16], lo_src[31:16]}
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(lo_src, hi_src, 0x7632);
#else
    return (hi_src & 0xffff0000u) | (lo_src >> 16);
#endif
}

BSW_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
BSW_HD int bsw_top_bit(uint32_t x)      // index of the highest set bit, x != 0
{
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}
BSW_HD int bsw_low_bit(uint32_t x)      // index of the lowest set bit, x != 0
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

constexpr int K1_S = TILE_LANES;            // stride (in words) between consecutive columns / words of one lane
constexpr int K1_QS_EXTRA = 8;              // query words per lane past nqw_max (one-hot planes need a zero block)
constexpr int K1_KEY_NONE = -1;

// 8 nibbles -> 8 bits: bit k set iff nibble k of w equals base b
BSW_HD uint32_t k1_eq8(uint32_t w, uint32_t b)
{
    uint32_t x = w ^ (b * 0x11111111u);
    x |= x >> 1;
    x |= x >> 2;
    uint32_t z = ~x & 0x11111111u;
    z = (z | (z >> 3)) & 0x03030303u;
    z = (z | (z >> 6)) & 0x000f000fu;
    z = (z | (z >> 12)) & 0xffu;
    return z;
}

// Score lookup for the matrix path: byte `nib` of the target base's matrix row (the RTL's 25:1 mux,
// sw_pe_array_mux_25to1_sel5_8_1.v:105-142), sign-extended.
BSW_HD int k1_lookup(uint32_t nib, uint32_t rlo, uint32_t rhi)
{
#if defined(__CUDA_ARCH__)
    return (int)(signed char)(__byte_perm(rlo, rhi, nib) & 0xffu);
#else
    const uint64_t both = (uint64_t)rlo | ((uint64_t)rhi << 32);
    return (int)(signed char)((both >> (8 * (nib & 7u))) & 0xffu);
#endif
}

// One extension.  eh/qs point at this lane's column 0 / word 0 (stride K1_S); tg at this lane's target word 0 in the
// tiled arena (stride K1_S).  eh must have qlen + 1 + K1_EH_SLACK columns; qs holds the tile's nqw_tile packed query
// words of this lane and has room for nqw_max + K1_QS_EXTRA words.
template <int VARIANT, int GENERIC, int SYM>
BSW_HD void k1_task(const DevParams& P, const int qlen, const int tlen, const int h0, const int w, const int nqw_tile,
                    uint32_t* eh, uint32_t* qs, const uint32_t* tg, SlotResult& res, const bool prep = true)
{
#define BSW_EHA(J) (eh + (J) * K1_S)
    constexpr bool ONEHOT = (VARIANT == 1 && GENERIC == 0);     // the branch-free path of the +a/-b scoring
    const int o_del = P.o_del, e_del = P.e_del, e_ins = P.e_ins;
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    const int zdrop = P.zdrop;
    const int mat = P.match, mis = -P.mismatch;
    // packed constants of the branch-free cell
    uint32_t c_mis = ((uint32_t)(GENERIC ? 0 : mis) << 16) | 0x8000u;        // {-b (or 0), -32768}
    uint32_t c_noe_del = (uint32_t)(-oe_del) << 16;                          // {-oe_del, 0}
    uint32_t c_noe_ins = (uint32_t)(-oe_ins) << 16;                          // {-oe_ins, 0}
    uint32_t c_ne_ins = (uint32_t)(-e_ins) << 16;                            // {-e_ins, 0}
    uint32_t c_eh = 0x80000000u | ((uint32_t)(-e_del) & 0xffffu);            // {-32768, -e_del}
    // V2 opens both gaps from M = M_old + s (not from h): {s - oe, -32768} added to {M_old + (a+b)*match, e}
    uint32_t c_mis_oed = ((uint32_t)((GENERIC ? 0 : mis) - oe_del) << 16) | 0x8000u;
    uint32_t c_mis_oei = ((uint32_t)((GENERIC ? 0 : mis) - oe_ins) << 16) | 0x8000u;
    uint32_t mul4[4];                                                        // V2 (query kept as nibbles): match bit 4k -> +(a+b) in the H half
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 4; ++k) mul4[k] = (uint32_t)(mat - mis) << (16 - 4 * k);
    uint32_t zero = P.zero;                                                  // 0, but opaque to ptxas (else it re-materialises a zero per cell)
    uint32_t mul[8];                                                         // (a+b) << (16-k)
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 8; ++k) mul[k] = (uint32_t)(mat - mis) << (16 - k);
#if defined(__CUDA_ARCH__)
    // keep the loop constants in ordinary registers (otherwise ptxas re-reads them from the constant bank per cell)
    asm volatile("" : "+r"(c_mis), "+r"(c_noe_del), "+r"(c_noe_ins), "+r"(c_ne_ins), "+r"(c_eh), "+r"(zero));
    asm volatile("" : "+r"(mul[0]), "+r"(mul[1]), "+r"(mul[2]), "+r"(mul[3]), "+r"(mul[4]), "+r"(mul[5]), "+r"(mul[6]), "+r"(mul[7]));
    if (VARIANT == 2) asm volatile("" : "+r"(c_mis_oed), "+r"(c_mis_oei), "+r"(mul4[0]), "+r"(mul4[1]), "+r"(mul4[2]), "+r"(mul4[3]));
#endif

    if (!prep) {
        // the query was prepared by an earlier call on the same lane (band retry of the fused seed kernel)
    } else if (ONEHOT) {
        // Re-code the query in place: per 32 columns, four words = the positions of A, C, G, T.  A chunk's match bits
        // for target base t are then one funnel shift of plane t (the RTL's mux_25to1 becomes a bit test).
        const int nblk = ((qlen + 31) >> 5) + 1;                 // one all-zero block past the end for the funnel shift
        for (int m = 0; m < nblk; ++m) {
            uint32_t wd[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int u = 0; u < 4; ++u) wd[u] = (4 * m + u < nqw_tile) ? qs[(4 * m + u) * K1_S] : 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (uint32_t b = 0; b < 4; ++b) {
                uint32_t pl = k1_eq8(wd[0], b) | (k1_eq8(wd[1], b) << 8) | (k1_eq8(wd[2], b) << 16) | (k1_eq8(wd[3], b) << 24);
                if (32 * m >= qlen) pl = 0;
                qs[(4 * m + (int)b) * K1_S] = pl;
            }
        }
    } else {
        qs[nqw_tile * K1_S] = 0;                                 // one zero word past the block for the funnel shift
    }

    // first row: eh[j].h = H(-1, j-1), all e = 0 (sx:1818; 1979,1957,1974; 1975-1978,1821)
    {
        eh[0] = (uint32_t)h0 << 16;
        int hv = h0 - P.o_ins;
        for (int j = 1; j <= qlen; ++j) { hv -= e_ins; eh[j * K1_S] = (uint32_t)imax(hv, 0) << 16; }
    }

    int max = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0;   // sx:889,1009,919,1019,1029,929
    int beg = 0, cend = qlen;            // cend = candidate end of the coming row (qlen, then previous end + 1)
    int resetmax = -1, stopmin = 0x7fffffff;
    uint32_t cells = 0;
    uint32_t tw = tg[0], tnext = 0;
    if (tlen > 8) tnext = tg[K1_S];

    for (int i = 0; i < tlen; ++i) {                                             // sx:1891
        if ((i & 7) == 0 && i) {
            tw = tnext;
            if (i + 8 < tlen) tnext = tg[((i >> 3) + 1) * K1_S];                 // prefetch the next 8 rows' bases
        }
        const uint32_t tb = tw & 15u;
        tw >>= 4;
        const uint32_t trep = tb * 0x11111111u;
        uint32_t rlo = 0, rhi = 0;
        if (GENERIC) { rlo = P.row_lo[tb]; rhi = P.row_hi[tb]; }
        const uint32_t* plane = qs + (tb & 3u) * K1_S;                           // ONEHOT: positions of this row's base

        int j0 = imax(beg, i - w);                                               // sx:1846,1894,1895,1803
        int lim = imin(imin(cend, i + w + 1), qlen);                             // sx:1980,1843,1897,1898,1842
        int fnz = 0x7fffffff, lnz = -1;                                          // V2 narrowing bookkeeping
        if (VARIANT == 1 && stopmin < j0) {
            // rare: the band clamp moved the start past mj+2; a zero in between ends the row (end' <= beg)
            const int zend = imin(j0, lim);
            for (int z = stopmin; z < zend; ++z)
                if ((*BSW_EHA(z) >> 16) == 0) { lim = imin(lim, z); break; }
        }
        int fc;                                                                  // first column (sx:1796,1795,1880,1835,849)
        if (VARIANT == 1 || j0 == 0) fc = imax(h0 - (o_del + e_del * (i + 1)), 0); else fc = 0;
        if (VARIANT == 1) {
            // trim the zero prefix (beg' = last zero + 1, sx:1766-1769) and the zero suffix (end' = first zero
            // >= mj+2, sx:1779,1782-1789) of the candidate window; interior zeros are caught chunk by chunk.
            while (j0 < lim && j0 <= resetmax && (*BSW_EHA(j0) >> 16) == 0) ++j0;
            while (lim > j0 && lim - 1 >= stopmin && (*BSW_EHA(lim - 1) >> 16) == 0) --lim;
        }
        uint32_t h1 = (uint32_t)fc << 16, f = 0;                                 // packed {value, 0}
        int b_eff = j0;
        int mkey = K1_KEY_NONE;
        int j = j0;
        uint32_t* ehp = BSW_EHA(j);

        // One DP cell on the branch-free path (see the header comment).  X = match bits (ONEHOT) or query nibbles (GENERIC).
#define BSW_K1_FAST(K, W, X, LIVE)                                                                   \
        {                                                                                            \
            uint32_t Wm;                                                                             \
            if (GENERIC) Wm = (uint32_t)k1_lookup(((X) >> (4 * (K))) & 15u, rlo, rhi) * 65536u + (W); \
            else         Wm = ((X) & (1u << (K))) * mul[K] + (W);                                    \
            const uint32_t hh = add_max_s16x2(Wm, c_mis, Wm << 16);                                  \
            const uint32_t h = max_s16x2(hh, f);                                                     \
            const uint32_t t = add_max_s16x2(h, c_noe_del, zero);                                    \
            const uint32_t nw = add_max_s16x2((W), c_eh, pack_hi_hi(h1, t));                         \
            if (SYM) f = add_max_s16x2(f, c_ne_ins, t);                                              \
            else     f = add_max_s16x2(f, c_ne_ins, add_max_s16x2(hh, c_noe_ins, zero));             \
            if (LIVE) { ehp[(K) * K1_S] = nw; ckey = (K) ? add_max((int)h, (K), ckey) : (int)h; h1 = h; } \
        }

        // The same cell under the V2 recurrence, for chunks whose H are all non-zero (the zero guard "M ? M + s : 0" is then
        // vacuous): both gap opens come from M = M_old + s, which is Wm plus a constant -- off the F chain.  XB = match bits at
        // bit 4k (XB2 = XB >> 16 for k >= 4) or the nibbles (GENERIC).  ACC collects "stored word != 0" per column: BWA's
        // narrowing scans for the first and last column whose h and e are not both zero.
#define BSW_K1_FAST2(K, W, XB, XB2)                                                                  \
        {                                                                                            \
            uint32_t Wm;                                                                             \
            if (GENERIC) Wm = (uint32_t)k1_lookup(((XB) >> (4 * (K))) & 15u, rlo, rhi) * 65536u + (W); \
            else         Wm = (((K) < 4 ? (XB) : (XB2)) & (1u << (4 * ((K) & 3)))) * mul4[(K) & 3] + (W); \
            const uint32_t hh = add_max_s16x2(Wm, c_mis, Wm << 16);                                  \
            const uint32_t h = max_s16x2(hh, f);                                                     \
            const uint32_t t = add_max_s16x2(Wm, c_mis_oed, zero);                                   \
            const uint32_t nw = add_max_s16x2((W), c_eh, pack_hi_hi(h1, t));                         \
            if (SYM) f = add_max_s16x2(f, c_ne_ins, t);                                              \
            else     f = add_max_s16x2(f, c_ne_ins, add_max_s16x2(Wm, c_mis_oei, zero));             \
            ehp[(K) * K1_S] = nw; ckey = (K) ? add_max((int)h, (K), ckey) : (int)h; h1 = h;          \
            acc += (nw != 0u ? 1u : 0u) << (K);                                                      \
        }

        // ... and for chunks that do hold a zero H: the guard "M ? M + s : 0" as one unsigned minimum.  mk+ = max(M_old + s, 0)
        // differs from the guarded value only where M_old == 0 and s > 0; (0 - (M_old << 16)) as an unsigned word is 0 there and
        // at least 0x80010000 everywhere else, so min_u32(mk+, that) is the guard.  (A negative M_old + s behaves like 0 in
        // everything that follows: h is a max with e, f >= 0 and both gap opens are clipped at 0.)
#define BSW_K1_GUARD2(K, W, XB, XB2)                                                                 \
        {                                                                                            \
            uint32_t Wm;                                                                             \
            if (GENERIC) Wm = (uint32_t)k1_lookup(((XB) >> (4 * (K))) & 15u, rlo, rhi) * 65536u + (W); \
            else         Wm = (((K) < 4 ? (XB) : (XB2)) & (1u << (4 * ((K) & 3)))) * mul4[(K) & 3] + (W); \
            const uint32_t mkp = add_max_s16x2(Wm, c_mis, zero);                                     \
            const uint32_t mk = umin32(mkp, 0u - ((W) & 0xffff0000u));                               \
            const uint32_t hh = max_s16x2(mk, Wm << 16);                                             \
            const uint32_t h = max_s16x2(hh, f);                                                     \
            const uint32_t t = add_max_s16x2(mk, c_noe_del, zero);                                   \
            const uint32_t nw = add_max_s16x2((W), c_eh, pack_hi_hi(h1, t));                         \
            if (SYM) f = add_max_s16x2(f, c_ne_ins, t);                                              \
            else     f = add_max_s16x2(f, c_ne_ins, add_max_s16x2(mk, c_noe_ins, zero));             \
            ehp[(K) * K1_S] = nw; ckey = (K) ? add_max((int)h, (K), ckey) : (int)h; h1 = h;          \
            acc += (nw != 0u ? 1u : 0u) << (K);                                                      \
        }

        bool stop = false;
        while (j < lim) {
            const int nv = lim - j;
            uint32_t x;                                      // ONEHOT: match bits of columns j..j+7; else: their nibbles
            if (ONEHOT) {
                const int m = j >> 5;
                x = funnel_r(plane[(4 * m) * K1_S], plane[(4 * m + 4) * K1_S], j & 31);
            } else {
                const int qi = j >> 3;
                x = funnel_r(qs[qi * K1_S], qs[(qi + 1) * K1_S], (j & 7) * 4);
            }
            if (VARIANT == 1 && ONEHOT && nv >= 16) {
                // two chunks under one head: one funnel shift yields the match bits of 16 columns, the 16 row-buffer
                // loads are issued together, and the zero test, the loop bookkeeping and the branches are paid once
                const uint32_t w0 = ehp[0 * K1_S], w1 = ehp[1 * K1_S], w2 = ehp[2 * K1_S], w3 = ehp[3 * K1_S];
                const uint32_t w4 = ehp[4 * K1_S], w5 = ehp[5 * K1_S], w6 = ehp[6 * K1_S], w7 = ehp[7 * K1_S];
                const uint32_t v0 = ehp[8 * K1_S], v1 = ehp[9 * K1_S], v2 = ehp[10 * K1_S], v3 = ehp[11 * K1_S];
                const uint32_t v4 = ehp[12 * K1_S], v5 = ehp[13 * K1_S], v6 = ehp[14 * K1_S], v7 = ehp[15 * K1_S];
                uint32_t zm = min3_u16x2(w0, w1, w2);
                zm = min3_u16x2(zm, w3, w4);
                zm = min3_u16x2(zm, w5, w6);
                zm = min3_u16x2(zm, w7, v0);
                zm = min3_u16x2(zm, v1, v2);
                zm = min3_u16x2(zm, v3, v4);
                zm = min3_u16x2(zm, v5, v6);
                zm = min3_u16x2(zm, v7, v7);
                if (zm >= 0x10000u) {                        // no zero H in the 16 columns
                    int ckey = K1_KEY_NONE;
                    BSW_K1_FAST(0, w0, x, true)
                    BSW_K1_FAST(1, w1, x, true)
                    BSW_K1_FAST(2, w2, x, true)
                    BSW_K1_FAST(3, w3, x, true)
                    BSW_K1_FAST(4, w4, x, true)
                    BSW_K1_FAST(5, w5, x, true)
                    BSW_K1_FAST(6, w6, x, true)
                    BSW_K1_FAST(7, w7, x, true)
                    mkey = imax(mkey, ckey + j);
                    ehp += 8 * K1_S;
                    const uint32_t x2 = x >> 8;
                    BSW_K1_FAST(0, v0, x2, true)
                    BSW_K1_FAST(1, v1, x2, true)
                    BSW_K1_FAST(2, v2, x2, true)
                    BSW_K1_FAST(3, v3, x2, true)
                    BSW_K1_FAST(4, v4, x2, true)
                    BSW_K1_FAST(5, v5, x2, true)
                    BSW_K1_FAST(6, v6, x2, true)
                    BSW_K1_FAST(7, v7, x2, true)
                    mkey = imax(mkey, ckey + j + 8);
                    j += 16; ehp += 8 * K1_S;
                    continue;
                }
            }
            if (VARIANT == 1) {
                const uint32_t w0 = ehp[0 * K1_S], w1 = ehp[1 * K1_S], w2 = ehp[2 * K1_S], w3 = ehp[3 * K1_S];
                const uint32_t w4 = ehp[4 * K1_S], w5 = ehp[5 * K1_S], w6 = ehp[6 * K1_S], w7 = ehp[7 * K1_S];
                int ckey = K1_KEY_NONE;                      // chunk-local key: (h << 16) + k
                if (nv >= 8) {
                    uint32_t zm = min3_u16x2(w0, w1, w2);
                    zm = min3_u16x2(zm, w3, w4);
                    zm = min3_u16x2(zm, w5, w6);
                    zm = min3_u16x2(zm, w7, w7);
                    if (zm >= 0x10000u) {                    // no zero H in the chunk
                        BSW_K1_FAST(0, w0, x, true)
                        BSW_K1_FAST(1, w1, x, true)
                        BSW_K1_FAST(2, w2, x, true)
                        BSW_K1_FAST(3, w3, x, true)
                        BSW_K1_FAST(4, w4, x, true)
                        BSW_K1_FAST(5, w5, x, true)
                        BSW_K1_FAST(6, w6, x, true)
                        BSW_K1_FAST(7, w7, x, true)
                        mkey = imax(mkey, ckey + j);
                        j += 8; ehp += 8 * K1_S;
                        continue;
                    }
                } else if (nv <= 4) {
                    // partial last chunk of at most four cells: half the work of the general one below
                    const uint32_t ones = 0xffffffffu;
                    uint32_t zm = min3_u16x2(w0, nv > 1 ? w1 : ones, nv > 2 ? w2 : ones);
                    zm = min3_u16x2(zm, nv > 3 ? w3 : ones, ones);
                    if (zm >= 0x10000u) {
                        BSW_K1_FAST(0, w0, x, true)
                        BSW_K1_FAST(1, w1, x, nv > 1)
                        BSW_K1_FAST(2, w2, x, nv > 2)
                        BSW_K1_FAST(3, w3, x, nv > 3)
                        mkey = imax(mkey, ckey + j);
                        j += nv; ehp += nv * K1_S;
                        break;
                    }
                } else {
                    // partial last chunk: the same code with the cells at or past lim made inert
                    const uint32_t ones = 0xffffffffu;
                    uint32_t zm = min3_u16x2(w0, nv > 1 ? w1 : ones, nv > 2 ? w2 : ones);
                    zm = min3_u16x2(zm, nv > 3 ? w3 : ones, nv > 4 ? w4 : ones);
                    zm = min3_u16x2(zm, nv > 5 ? w5 : ones, nv > 6 ? w6 : ones);
                    if (zm >= 0x10000u) {
                        BSW_K1_FAST(0, w0, x, true)
                        BSW_K1_FAST(1, w1, x, nv > 1)
                        BSW_K1_FAST(2, w2, x, nv > 2)
                        BSW_K1_FAST(3, w3, x, nv > 3)
                        BSW_K1_FAST(4, w4, x, nv > 4)
                        BSW_K1_FAST(5, w5, x, nv > 5)
                        BSW_K1_FAST(6, w6, x, nv > 6)
                        mkey = imax(mkey, ckey + j);
                        j += nv; ehp += nv * K1_S;
                        break;
                    }
                }
            }
            if (VARIANT == 2 && nv >= 8) {
                const uint32_t w0 = ehp[0 * K1_S], w1 = ehp[1 * K1_S], w2 = ehp[2 * K1_S], w3 = ehp[3 * K1_S];
                const uint32_t w4 = ehp[4 * K1_S], w5 = ehp[5 * K1_S], w6 = ehp[6 * K1_S], w7 = ehp[7 * K1_S];
                uint32_t zm = min3_u16x2(w0, w1, w2);
                zm = min3_u16x2(zm, w3, w4);
                zm = min3_u16x2(zm, w5, w6);
                zm = min3_u16x2(zm, w7, w7);
                uint32_t xb = x, xb2 = 0;
                if (!GENERIC) {
                    uint32_t y = x ^ trep;                   // nibble == 0 <-> match
                    y |= y >> 1; y |= y >> 2;
                    xb = ~y & 0x11111111u; xb2 = xb >> 16;
                }
                int ckey = K1_KEY_NONE;
                uint32_t acc = 0;
                if (zm >= 0x10000u) {                        // no zero H in the chunk: the zero guard never fires
                    BSW_K1_FAST2(0, w0, xb, xb2)
                    BSW_K1_FAST2(1, w1, xb, xb2)
                    BSW_K1_FAST2(2, w2, xb, xb2)
                    BSW_K1_FAST2(3, w3, xb, xb2)
                    BSW_K1_FAST2(4, w4, xb, xb2)
                    BSW_K1_FAST2(5, w5, xb, xb2)
                    BSW_K1_FAST2(6, w6, xb, xb2)
                    BSW_K1_FAST2(7, w7, xb, xb2)
                } else {
                    BSW_K1_GUARD2(0, w0, xb, xb2)
                    BSW_K1_GUARD2(1, w1, xb, xb2)
                    BSW_K1_GUARD2(2, w2, xb, xb2)
                    BSW_K1_GUARD2(3, w3, xb, xb2)
                    BSW_K1_GUARD2(4, w4, xb, xb2)
                    BSW_K1_GUARD2(5, w5, xb, xb2)
                    BSW_K1_GUARD2(6, w6, xb, xb2)
                    BSW_K1_GUARD2(7, w7, xb, xb2)
                }
                mkey = imax(mkey, ckey + j);
                if (acc) { lnz = j + bsw_top_bit(acc); fnz = imin(fnz, j + bsw_low_bit(acc)); }
                j += 8; ehp += 8 * K1_S;
                continue;
            }
            // careful path: cell by cell, with the narrowing events (V1) or the V2 recurrence
            const int kmax = nv < 8 ? nv : 8;
            for (int k = 0; k < kmax; ++k, ++j, ehp += K1_S) {
                const uint32_t wd = *ehp;
                int M = (int)(wd >> 16), e = (int)(wd & 0xffffu);
                if (VARIANT == 1 && M == 0) {
                    if (j <= resetmax) { f = 0; h1 = (uint32_t)fc << 16; mkey = K1_KEY_NONE; b_eff = j + 1; continue; }   // beg' = j+1
                    if (j >= stopmin) { lim = j; stop = true; break; }                                  // end' = j
                }
                int s;
                if (ONEHOT) s = ((x >> k) & 1u) ? mat : mis;
                else if (GENERIC) s = k1_lookup((x >> (4 * k)) & 15u, rlo, rhi);
                else s = (((x >> (4 * k)) ^ trep) & 15u) ? mis : mat;
                const int fi = (int)(f >> 16), h1i = (int)(h1 >> 16);
                int h, g;
                if (VARIANT == 1) { h = imax(imax(M + s, e), fi); g = h; }                               // sx:1797,1798,1809
                else { M = M ? M + s : 0; h = imax(imax(M, e), fi); g = M; }                             // upstream BWA
                mkey = imax(mkey, h * 65536 + j);                                                        // sx:1808,1816
                int t = imax(g - oe_del, 0);                                                             // sx:1866,1862
                e = imax(e - e_del, t);                                                                  // sx:1770-1771
                t = imax(g - oe_ins, 0);                                                                 // sx:1863,1865
                f = (uint32_t)imax(fi - e_ins, t) << 16;                                                 // sx:1780-1781
                if (VARIANT == 2) { if ((h1i | e) != 0) { lnz = j; fnz = imin(fnz, j); } }
                *ehp = ((uint32_t)h1i << 16) | (uint32_t)e;                                              // sx:1776
                h1 = (uint32_t)h << 16;
            }
            if (stop) break;
        }
#undef BSW_K1_FAST
#undef BSW_K1_FAST2
#undef BSW_K1_GUARD2

        const int e_eff = lim;
        const int h1v = (int)(h1 >> 16);
        if (e_eff > b_eff) cells += (uint32_t)(e_eff - b_eff);
        *BSW_EHA(e_eff) = h1;                                                    // eh[end] = {h1, e=0}: sx:1775,1904
        if (VARIANT == 2) { if (h1v != 0) lnz = e_eff; }
        const int j_after = e_eff > b_eff ? e_eff : b_eff;                       // value of j after the reference's loop
        if (j_after == qlen) {                                                   // sx:1768,1913
            if (!(gscore > h1v)) { max_ie = i; gscore = h1v; }                   // sx:1941,1829,1831
        }
        int m, mj;
        if (mkey < 0) { m = 0; mj = -1; } else { m = mkey >> 16; mj = mkey & 0xffff; }
        if (m == 0) break;                                                       // sx:1942,1686-1687
        if (m > max) {                                                           // sx:1959
            max = m; max_i = i; max_j = mj;                                      // sx:1810,1833,1801
            const int d = mj > i ? mj - i : i - mj;
            max_off = max_off > d ? max_off : d;                                 // sx:1845,1707-1708,1964,1812
        } else if (zdrop > 0) {                                                  // ksw_extend2 z-drop (not in the RTL)
            const int di = i - max_i, dj = mj - max_j;
            if (di > dj) { if (max - m - (di - dj) * e_del > zdrop) break; }
            else         { if (max - m - (dj - di) * e_ins > zdrop) break; }
        }
        if (VARIANT == 1) {
            beg = b_eff; cend = e_eff + 1; resetmax = mj; stopmin = mj + 2;      // lazy form of sx:1766-1769,1779,1782-1789
        } else {
            // upstream BWA: drop leading/trailing columns whose h and e are both zero
            const int nb = fnz < e_eff ? fnz : e_eff;
            const int jl = lnz > nb - 1 ? lnz : nb - 1;
            beg = nb; cend = imin(jl + 2, qlen);
        }
    }
    res.score = max; res.qle = max_j + 1; res.tle = max_i + 1; res.gtle = max_ie + 1;     // sx:1315-1375,1841,1868,1794
    res.gscore = gscore; res.max_off = max_off; res.cells = (int32_t)cells; res.status = STATUS_OK;   // sx:1792,1815
#undef BSW_EHA
}

}  // namespace bsw
