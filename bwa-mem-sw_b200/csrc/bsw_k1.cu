// bsw_k1.cu -- K1: inter-task extension kernel, one thread per extension task (sm_100a).
//
// Replaces the 80 free-running PEs of the reference (sw_pe_array.v:1133-1494, one task per PE,
// sw_pe_array_sw_extend.v FSM :1639-1705) by thousands of resident threads, each walking one
// banded affine-gap extension row by row exactly as the RTL / ksw_extend2 does.
//
// Per-thread state
//   * the rolling row buffer eh[0..qlen] (BWA's eh_t, RTL's eh_arr: sw_pe_array_sw_extend_eh_arr.v)
//     lives in shared memory, one 32-bit word {E[31:16], H[15:0]} per column (the RTL packs
//     {E[15:8],H[7:0]}), laid out eh[j][lane] so that a warp's 32 tasks hit 32 different banks;
//     WIDE instantiations keep {H,E} as two int32 (uint2) for tasks whose score bound exceeds int16.
//   * the query, 4 bit per base, also in shared memory (qs[word][lane]); 8 columns per LDS.
//   * the target is read from HBM one 32-bit word (8 rows) at a time.
// Scoring: FAST (matrix is +a / -b and the task holds no N) compares nibbles by XOR; GENERIC looks the
// score up in the target base's matrix row with PRMT (any int8 5x5 matrix, N included) -- the RTL's
// 25:1 mux (sw_pe_array_mux_25to1_sel5_8_1.v).
//
// Band narrowing (V1).  The reference recomputes [beg,end) after every row by scanning the stored row
// for the run of non-zero H around mj (sw_pe_array_sw_extend.v:1766-1769,1779,1782-1789).  Scanning
// costs a second pass.  K1 evaluates the same rule lazily inside the NEXT row, which reads every
// eh[j].h of the candidate window anyway: a zero at j <= mj restarts the row at j+1 (that is beg' = last
// zero + 1), a zero at j >= mj+2 ends it (that is end').  The result is bit-identical; see DESIGN.md.
#include "bsw_device.cuh"
#include "bsw_kernels.h"

namespace bsw {

constexpr int K1_NT = 32;   // one warp per CTA: no block-level synchronisation anywhere in K1

__device__ __forceinline__ int imax(int a, int b) { return a > b ? a : b; }
__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }

template <int WIDE> struct EhWord;
template <> struct EhWord<0> {
    typedef uint32_t type;
    static __device__ __forceinline__ void unpack(type w, int& h, int& e) { h = (int)(w & 0xffffu); e = (int)(w >> 16); }
    static __device__ __forceinline__ type pack(int h, int e) { return (uint32_t)h | ((uint32_t)e << 16); }
};
template <> struct EhWord<1> {
    typedef uint2 type;
    static __device__ __forceinline__ void unpack(type w, int& h, int& e) { h = (int)w.x; e = (int)w.y; }
    static __device__ __forceinline__ type pack(int h, int e) { return make_uint2((uint32_t)h, (uint32_t)e); }
};

template <int WIDE> struct MaxKey;
template <> struct MaxKey<0> {       // h < 32768, j < 65536: (h<<16 | j); max() keeps the right-most arg-max
    typedef int type;
    static __device__ __forceinline__ type none() { return -1; }
    static __device__ __forceinline__ type make(int h, int j) { return (h << 16) | j; }
    static __device__ __forceinline__ void split(type k, int& m, int& mj) { if (k < 0) { m = 0; mj = -1; } else { m = k >> 16; mj = k & 0xffff; } }
};
template <> struct MaxKey<1> {
    typedef long long type;
    static __device__ __forceinline__ type none() { return -1; }
    static __device__ __forceinline__ type make(int h, int j) { return ((long long)h << 32) | (unsigned)j; }
    static __device__ __forceinline__ void split(type k, int& m, int& mj) { if (k < 0) { m = 0; mj = -1; } else { m = (int)(k >> 32); mj = (int)(k & 0xffffffffll); } }
};

// One DP cell at column j+K.  Mirrors the RTL datapath (sw_pe_array_sw_extend.v):
//   M=eh.h, e=eh.e (:1799,1772) ; eh.h=h1 (:1776) ; h=M+s (:1797) ; h=max(h,e) (:1798) ; h=max(h,f) (:1809)
//   m/mj (:1808,1816) ; t=max(0,h-oe_del) (:1866,1862) ; e=max(e-e_del,t) (:1770-1771)
//   t=max(0,h-oe_ins) (:1863,1865) ; f=max(f-e_ins,t) (:1780-1781)
#define BSW_K1_CELL(K, NIB)                                                                         \
    {                                                                                               \
        const int jj = j + (K);                                                                     \
        const typename EH::type wd = ehp[(K) * K1_NT];                                              \
        int M, e;                                                                                   \
        EH::unpack(wd, M, e);                                                                       \
        bool skip = false;                                                                          \
        if (VARIANT == 1) {                                                                         \
            if (__builtin_expect(M == 0, 0)) {                                                      \
                if (jj <= resetmax) { f = 0; h1 = fc; mkey = MK::none(); b_eff = jj + 1; skip = true; } \
                else if (jj >= stopmin) { lim = jj; goto row_done; }                                 \
            }                                                                                       \
        }                                                                                           \
        if (!skip) {                                                                                \
            int s;                                                                                  \
            if (GENERIC) { s = (int)(signed char)(__byte_perm(rlo, rhi, (NIB)) & 0xffu); }          \
            else         { s = (NIB) ? mis : mat; }                                                 \
            int h, g;                                                                               \
            if (VARIANT == 1) { h = imax(imax(M + s, e), f); g = h; }                                \
            else { M = M ? M + s : 0; h = imax(imax(M, e), f); g = M; }                              \
            mkey = mkey > MK::make(h, jj) ? mkey : MK::make(h, jj);                                  \
            int t = imax(g - oe_del, 0);                                                            \
            e = imax(e - e_del, t);                                                                 \
            if (!SYM) t = imax(g - oe_ins, 0);                                                      \
            f = imax(f - e_ins, t);                                                                 \
            if (VARIANT == 2) { if ((h1 | e) != 0) { lnz = jj; fnz = imin(fnz, jj); } }              \
            ehp[(K) * K1_NT] = EH::pack(h1, e);                                                     \
            h1 = h;                                                                                 \
        }                                                                                           \
    }

// Branch-free cell for chunks that hold no zero H (V1, 16-bit row buffer).  The E update and the
// re-packing of {E,H} are one packed DPX op: VIADDMNMX.S16x2({e,M} + {-e_del,-32768}, {t,h1}) = {max(e-e_del,t), h1}.
#define BSW_K1_FAST(K, W, NIB)                                                                      \
    {                                                                                               \
        const int M = (int)((W) & 0xffffu), e = (int)((W) >> 16);                                   \
        int s;                                                                                      \
        if (GENERIC) { s = (int)(signed char)(__byte_perm(rlo, rhi, (NIB)) & 0xffu); }              \
        else         { s = (NIB) ? mis : mat; }                                                     \
        const int h = imax(__viaddmax_s32(M, s, e), f);                                             \
        const int t = __viaddmax_s32_relu(h, noe_del, 0);                                           \
        ehp[(K) * K1_NT] = __viaddmax_s16x2((W), ce_pack, ((uint32_t)t << 16) + (uint32_t)h1);      \
        if (SYM) f = __viaddmax_s32(f, ne_ins, t);                                                  \
        else     f = __viaddmax_s32(f, ne_ins, __viaddmax_s32_relu(h, noe_ins, 0));                 \
        mkey = imax(mkey, (h << 16) + (j + (K)));                                                   \
        h1 = h;                                                                                     \
    }

template <int VARIANT, int GENERIC, int SYM, int WIDE>
__global__ void __launch_bounds__(K1_NT) k1_extend_kernel(const LaunchArgs A)
{
    typedef EhWord<WIDE> EH;
    typedef MaxKey<WIDE> MK;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const uint32_t slot = A.slot0 + blockIdx.x * K1_NT + lane;
    unsigned long long my_cells = 0;
    if (slot < A.slot1) {
        const uint32_t task = A.order[slot];
        const int qlen = A.t.qlen[task], tlen = A.t.tlen[task], h0 = A.t.h0[task], w = A.t.w[task];
        const uint32_t* __restrict__ qg = A.t.qseq + A.t.qoffw[task];
        const uint32_t* __restrict__ tg = A.t.tseq + A.t.toffw[task];
        const int o_del = A.p.o_del, e_del = A.p.e_del, e_ins = A.p.e_ins;
        const int oe_del = A.p.o_del + A.p.e_del, oe_ins = A.p.o_ins + A.p.e_ins;
        const int zdrop = A.p.zdrop;
        int mat = A.p.match, mis = -A.p.mismatch;
        int noe_del = -oe_del, noe_ins = -oe_ins, ne_ins = -e_ins;
        uint32_t ce_pack = 0x8000u | ((uint32_t)(-e_del) << 16);
        // keep the loop constants in ordinary registers (otherwise ptxas re-loads them with LDCU per cell)
        asm volatile("" : "+r"(mat), "+r"(mis), "+r"(noe_del), "+r"(noe_ins), "+r"(ne_ins), "+r"(ce_pack));

        typename EH::type* eh = reinterpret_cast<typename EH::type*>(smem_raw) + lane;          // eh[j * K1_NT]
        uint32_t* qs = reinterpret_cast<uint32_t*>(reinterpret_cast<typename EH::type*>(smem_raw) + (size_t)(A.qmax + 1) * K1_NT) + lane;

        // stage the query (8 bases per word) and fill the first row (sx:1818,1979,1975-1978)
        {
            const int nq = (qlen + 7) >> 3;
            for (int k = 0; k < nq; k += 4) {
                const uint4 v = *reinterpret_cast<const uint4*>(qg + k);
                qs[(k + 0) * K1_NT] = v.x; qs[(k + 1) * K1_NT] = v.y; qs[(k + 2) * K1_NT] = v.z; qs[(k + 3) * K1_NT] = v.w;
            }
            qs[(((nq + 3) & ~3)) * K1_NT] = 0;      // one word of slack for the funnel shift below
            eh[0] = EH::pack(h0, 0);
            int hv = h0 - A.p.o_ins;
            for (int j = 1; j <= qlen; ++j) { hv -= e_ins; eh[j * K1_NT] = EH::pack(imax(hv, 0), 0); }
        }

        int max = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0;   // sx:889,1009,919,1019,1029,929
        int beg = 0, cend = qlen;            // cend = candidate end for the coming row (qlen, then e_eff+1)
        int resetmax = -1, stopmin = 0x7fffffff;
        uint32_t cells = 0;
        uint32_t tw = tg[0];

        for (int i = 0; i < tlen; ++i) {                                         // sx:1891
            if ((i & 7) == 0 && i) tw = tg[i >> 3];
            const uint32_t tb = (tw >> ((i & 7) * 4)) & 15u;
            const uint32_t trep = tb * 0x11111111u;
            uint32_t rlo = 0, rhi = 0;
            if (GENERIC) { rlo = A.p.row_lo[tb]; rhi = A.p.row_hi[tb]; }

            int j0 = imax(beg, i - w);                                           // sx:1846,1894,1895,1803
            int lim = imin(imin(cend, i + w + 1), qlen);                         // sx:1980,1843,1897,1898,1842
            int fnz = 0x7fffffff, lnz = -1;                                      // V2 narrowing bookkeeping
            if (VARIANT == 1 && stopmin < j0) {
                // rare: the band clamp moved the start past mj+2; a zero in between ends the row (end' <= beg')
                const int zend = imin(j0, lim);
                for (int z = stopmin; z < zend; ++z) {
                    int M, e; EH::unpack(eh[z * K1_NT], M, e);
                    if (M == 0) { lim = imin(lim, z); break; }
                }
            }
            int fc;                                                              // first column (sx:1796,1795,1880,1835,849)
            if (VARIANT == 1 || j0 == 0) fc = imax(h0 - (o_del + e_del * (i + 1)), 0); else fc = 0;
            if (VARIANT == 1) {
                // trim the zero prefix (beg' = last zero + 1, sx:1766-1769) and the zero suffix
                // (end' = first zero >= mj+2, sx:1779,1782-1789) of the candidate window; interior zeros,
                // which are rare, are caught cell by cell below.
                while (j0 < lim && j0 <= resetmax) {
                    int M, e; EH::unpack(eh[j0 * K1_NT], M, e);
                    if (M) break;
                    ++j0;
                }
                while (lim > j0 && lim - 1 >= stopmin) {
                    int M, e; EH::unpack(eh[(lim - 1) * K1_NT], M, e);
                    if (M) break;
                    --lim;
                }
            }
            int h1 = fc, f = 0, b_eff = j0;
            typename MK::type mkey = MK::none();
            int j = j0;
            {
                typename EH::type* ehp = eh + j * K1_NT;
                // 8-column chunks: one aligned view of the packed query per chunk
                while (j + 8 <= lim) {
                    const int qi = j >> 3, sh = (j & 7) * 4;
                    const uint32_t qa = __funnelshift_r(qs[qi * K1_NT], qs[(qi + 1) * K1_NT], sh);
                    const uint32_t x = GENERIC ? qa : (qa ^ trep);
                    if constexpr (VARIANT == 1 && WIDE == 0) {
                        const uint32_t w0 = ehp[0 * K1_NT], w1 = ehp[1 * K1_NT], w2 = ehp[2 * K1_NT], w3 = ehp[3 * K1_NT];
                        const uint32_t w4 = ehp[4 * K1_NT], w5 = ehp[5 * K1_NT], w6 = ehp[6 * K1_NT], w7 = ehp[7 * K1_NT];
                        uint32_t zm = __vimin3_u16x2(w0, w1, w2);
                        zm = __vimin3_u16x2(zm, w3, w4);
                        zm = __vimin3_u16x2(zm, w5, w6);
                        zm = __vminu2(zm, w7);
                        if (__builtin_expect((zm & 0xffffu) != 0, 1)) {
                            BSW_K1_FAST(0, w0, GENERIC ? (x & 15u) : (x & 0x0000000fu))
                            BSW_K1_FAST(1, w1, GENERIC ? ((x >> 4) & 15u) : (x & 0x000000f0u))
                            BSW_K1_FAST(2, w2, GENERIC ? ((x >> 8) & 15u) : (x & 0x00000f00u))
                            BSW_K1_FAST(3, w3, GENERIC ? ((x >> 12) & 15u) : (x & 0x0000f000u))
                            BSW_K1_FAST(4, w4, GENERIC ? ((x >> 16) & 15u) : (x & 0x000f0000u))
                            BSW_K1_FAST(5, w5, GENERIC ? ((x >> 20) & 15u) : (x & 0x00f00000u))
                            BSW_K1_FAST(6, w6, GENERIC ? ((x >> 24) & 15u) : (x & 0x0f000000u))
                            BSW_K1_FAST(7, w7, GENERIC ? ((x >> 28) & 15u) : (x & 0xf0000000u))
                            j += 8; ehp += 8 * K1_NT;
                            continue;
                        }
                    }
                    BSW_K1_CELL(0, GENERIC ? (x & 15u) : (x & 0x0000000fu))
                    BSW_K1_CELL(1, GENERIC ? ((x >> 4) & 15u) : (x & 0x000000f0u))
                    BSW_K1_CELL(2, GENERIC ? ((x >> 8) & 15u) : (x & 0x00000f00u))
                    BSW_K1_CELL(3, GENERIC ? ((x >> 12) & 15u) : (x & 0x0000f000u))
                    BSW_K1_CELL(4, GENERIC ? ((x >> 16) & 15u) : (x & 0x000f0000u))
                    BSW_K1_CELL(5, GENERIC ? ((x >> 20) & 15u) : (x & 0x00f00000u))
                    BSW_K1_CELL(6, GENERIC ? ((x >> 24) & 15u) : (x & 0x0f000000u))
                    BSW_K1_CELL(7, GENERIC ? ((x >> 28) & 15u) : (x & 0xf0000000u))
                    j += 8; ehp += 8 * K1_NT;
                }
                if (j < lim) {
                    const int qi = j >> 3, sh = (j & 7) * 4;
                    uint32_t qa = __funnelshift_r(qs[qi * K1_NT], qs[(qi + 1) * K1_NT], sh);
                    uint32_t x = GENERIC ? qa : (qa ^ trep);
                    while (j < lim) {
                        BSW_K1_CELL(0, (x & 15u))
                        x >>= 4; ++j; ehp += K1_NT;
                    }
                }
            }
        row_done:
            const int e_eff = lim;
            if (e_eff > b_eff) cells += (uint32_t)(e_eff - b_eff);
            eh[e_eff * K1_NT] = EH::pack(h1, 0);                                  // sx:1775,1904
            if (VARIANT == 2) { if (h1 != 0) lnz = e_eff; }
            const int j_after = e_eff > b_eff ? e_eff : b_eff;
            if (j_after == qlen) {                                               // sx:1768,1913
                if (!(gscore > h1)) { max_ie = i; gscore = h1; }                 // sx:1941,1829,1831
            }
            int m, mj;
            MK::split(mkey, m, mj);
            if (m == 0) break;                                                   // sx:1942
            if (m > max) {                                                       // sx:1959
                max = m; max_i = i; max_j = mj;
                const int d = mj > i ? mj - i : i - mj;
                max_off = max_off > d ? max_off : d;                             // sx:1707-1708,1812
            } else if (zdrop > 0) {                                              // ksw_extend2 z-drop (not in the RTL)
                const int di = i - max_i, dj = mj - max_j;
                if (di > dj) { if (max - m - (di - dj) * e_del > zdrop) break; }
                else         { if (max - m - (dj - di) * e_ins > zdrop) break; }
            }
            if (VARIANT == 1) {
                beg = b_eff; cend = e_eff + 1; resetmax = mj; stopmin = mj + 2;   // lazy form of sx:1766-1769,1779,1782-1789
            } else {
                // upstream BWA: drop leading/trailing columns whose h and e are both zero
                const int nb = fnz < e_eff ? fnz : e_eff;
                const int jl = lnz > nb - 1 ? lnz : nb - 1;
                beg = nb; cend = imin(jl + 2, qlen);
            }
        }
        int4* o = A.t.out + 2 * (size_t)task;
        o[0] = make_int4(max, max_j + 1, max_i + 1, max_ie + 1);                  // sx:1315-1375 (score,qle,tle,gtle)
        o[1] = make_int4(gscore, max_off, (int)cells, STATUS_OK);                 //              (gscore,max_off) + cells
        my_cells = cells;
    }
    // one atomic per warp for the device-side cell counter
    for (int o = 16; o; o >>= 1) my_cells += __shfl_xor_sync(0xffffffffu, my_cells, o);
    if (lane == 0 && A.cells_total && my_cells) atomicAdd(A.cells_total, my_cells);
}

size_t k1_smem_bytes(int qmax, int wide)
{
    const size_t ehb = wide ? 8 : 4;
    const size_t qwords = (size_t)(((qmax + 7) >> 3) + 3 & ~3) + 1;
    return ((size_t)(qmax + 1) * ehb + qwords * 4) * K1_NT;
}

template <int VARIANT, int GENERIC, int SYM, int WIDE>
static cudaError_t k1_launch_t(const LaunchArgs& a, cudaStream_t st)
{
    const uint32_t n = a.slot1 - a.slot0;
    if (!n) return cudaSuccess;
    const size_t smem = k1_smem_bytes(a.qmax, WIDE);
    auto kern = k1_extend_kernel<VARIANT, GENERIC, SYM, WIDE>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    kern<<<(n + K1_NT - 1) / K1_NT, K1_NT, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t k1_launch(const LaunchArgs& a, int variant, int generic, int sym, int wide, cudaStream_t st)
{
#define BSW_K1_DISPATCH(V, G, S, W) if (variant == V && generic == G && sym == S && wide == W) return k1_launch_t<V, G, S, W>(a, st);
    BSW_K1_DISPATCH(1, 0, 1, 0) BSW_K1_DISPATCH(1, 0, 0, 0) BSW_K1_DISPATCH(1, 1, 1, 0) BSW_K1_DISPATCH(1, 1, 0, 0)
    BSW_K1_DISPATCH(1, 0, 1, 1) BSW_K1_DISPATCH(1, 0, 0, 1) BSW_K1_DISPATCH(1, 1, 1, 1) BSW_K1_DISPATCH(1, 1, 0, 1)
    BSW_K1_DISPATCH(2, 0, 1, 0) BSW_K1_DISPATCH(2, 0, 0, 0) BSW_K1_DISPATCH(2, 1, 1, 0) BSW_K1_DISPATCH(2, 1, 0, 0)
    BSW_K1_DISPATCH(2, 0, 1, 1) BSW_K1_DISPATCH(2, 0, 0, 1) BSW_K1_DISPATCH(2, 1, 1, 1) BSW_K1_DISPATCH(2, 1, 0, 1)
#undef BSW_K1_DISPATCH
    return cudaErrorInvalidValue;
}

}  // namespace bsw
