// bsw_k2s.cu -- K2S: intra-task extension kernel for long tasks with NARROW live windows: SW = 8 lanes per task,
// four tasks per warp (sm_100a).
//
// Same row-parallel algorithm as K2 (bsw_k2.cu: per-lane 8 columns with zero carry-in, prefix-max of the F carries by
// __shfl_up_sync, second pass, arg-max key by REDUX, narrowing as two masked bit scans), but a group is 8 lanes x 8
// columns = 64 columns.  With PacBio-like error rates the narrowed window of a 1-10 kb extension averages ~50 columns
// (max a few hundred), so a 256-column K2 group leaves 26 of 32 lanes idle; here the four 8-lane sub-warps of a warp
// carry four independent tasks, every __shfl / REDUX / __syncwarp is issued with the sub-warp's own member mask, and a
// wide row is simply more 64-column rounds of the same sub-warp.  The row buffer is the 2 048-column ring of K2
// (64-aligned groups never wrap), the packed query and target are read from the source arena.
#include <cuda_runtime.h>
#include "bsw_device.cuh"
#include "bsw_k1_core.cuh"
#include "bsw_kernels.h"

namespace bsw {

constexpr int K2S_RING_MAX = 2048;
constexpr int K2S_HDR_BYTES = 128;

template <int GENERIC>
__device__ __forceinline__ int k2s_score(uint32_t nib_or_xor, int mat, int mis, uint32_t rlo, uint32_t rhi)
{
    if (GENERIC) return k1_lookup(nib_or_xor, rlo, rhi);
    return nib_or_xor ? mis : mat;
}

__device__ __forceinline__ uint32_t k2s_range_mask(int base, int a, int b)
{
    const int lo = imax(a - base, 0), hi = imin(b - base, 31);
    if (lo > hi) return 0u;
    return (0xffffffffu << lo) & (0xffffffffu >> (31 - hi));
}

// per-task shared memory: zero bits (ring/32 words) + row ring (ring + 8 words), 16-byte aligned
__host__ __device__ inline size_t k2s_task_words(int rcap) { return (size_t)((rcap >> 5) + 4 + rcap + 8 + 4 + 3) & ~(size_t)3; }

template <int GENERIC, int SW>
__global__ void __launch_bounds__(32) k2s_extend_kernel(const __grid_constant__ LaunchArgs A)
{
    constexpr int TPW = 32 / SW;                       // tasks per warp
    constexpr int GROUP = 8 * SW;                      // columns per round
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x, sub = lane / SW, sl = lane % SW;
    const unsigned smask = (SW == 32) ? 0xffffffffu : (((1u << SW) - 1u) << (sub * SW));
    const uint32_t tile_raw = blockIdx.x * TPW + sub;
    const bool has_task = tile_raw < A.ntiles;         // a sub-warp without a task idles through the row loop (the warp
    const uint32_t tile = has_task ? tile_raw : A.ntiles - 1;   // re-converges once per row, see below)
    const TileHdr hd = A.tiles[tile];
    const uint32_t slot = hd.slot0;
    const SlotParam sp = A.slots[slot];
    const int qlen = sp.qlen, tlen = sp.tlen, h0 = sp.h0, w = sp.w;
    // Row buffer: a ring of A.ring_cols columns (power of two, chosen by the host so that the first row fits).  A row whose
    // 64-aligned span does not fit ends the task with STATUS_OVERFLOW and the host reruns it on K2.
    const int rcap = A.ring_cols;
    const int rm = rcap - 1;
    const int nqw = (qlen + 7) >> 3;

    uint32_t* base = reinterpret_cast<uint32_t*>(smem_raw + K2S_HDR_BYTES) + (size_t)sub * k2s_task_words(rcap);
    uint32_t* zb = base;                                                          // rcap/32 words of zero bits
    uint32_t* eh = reinterpret_cast<uint32_t*>((reinterpret_cast<uintptr_t>(zb + (rcap >> 5) + 4) + 15) & ~(uintptr_t)15);
    const uint32_t* qs = A.arena + (size_t)hd.qoff16 * 4u;                        // packed query, source arena
    const uint32_t* tg = A.arena + (size_t)hd.toff16 * 4u;

    const int o_del = A.p.o_del, e_del = A.p.e_del, e_ins = A.p.e_ins;
    const int oe_del = A.p.o_del + A.p.e_del, oe_ins = A.p.o_ins + A.p.e_ins;
    const int zdrop = A.p.zdrop;
    const int mat = A.p.match, mis = -A.p.mismatch;
    const uint32_t ce_pack = 0x8000u | ((uint32_t)(-e_del) << 16);
    const int e8 = 8 * e_ins, eg = GROUP * e_ins;

    // first row: eh[j].h = H(-1, j-1), all e = 0 (sx:1818; 1979,1957,1974; 1975-1978,1821)
    for (int j = sl; j < rcap + 8; j += SW) {
        int hv = (j == 0) ? h0 : imax(h0 - A.p.o_ins - j * e_ins, 0);
        if (j > qlen) hv = 0;
        eh[j] = (uint32_t)hv;
    }
    __syncwarp();

    int max = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0;   // sx:889,1009,919,1019,1029,929
    int beg = 0, end = qlen;                                                       // sx:769,779
    unsigned long long cells = 0;
    uint32_t tw = 0;

    // Every lane of a sub-warp carries the same row-level scalars, so the branches below are sub-warp-uniform.  The four
    // sub-warps are re-converged with a full-warp barrier once per row: without it they drift apart and the hardware
    // issues every instruction four times, 8 lanes at a time.
    bool done = !has_task, overflow = false;
    const int tmaxw = __reduce_max_sync(0xffffffffu, has_task ? tlen : 0);
    for (int i = 0; i < tmaxw; ++i) {                                              // sx:1891
      if (__all_sync(0xffffffffu, done || i >= tlen)) break;
      if (!done && i < tlen) do {
        if ((i & 7) == 0) tw = __ldg(tg + (i >> 3));
        const uint32_t tb = (tw >> ((i & 7) * 4)) & 15u;
        const uint32_t trep = tb * 0x11111111u;
        uint32_t rlo = 0, rhi = 0;
        if (GENERIC) { rlo = A.p.row_lo[tb]; rhi = A.p.row_hi[tb]; }

        const int j0 = imax(beg, i - w);                                           // sx:1846,1894,1895,1803
        const int lim = imin(imin(end, i + w + 1), qlen);                          // sx:1980,1843,1897,1898,1842
        const int fc = imax(h0 - (o_del + e_del * (i + 1)), 0);                    // V1: unconditional (sx:1796,1795,1880,1835,849)
        if (lim <= j0) {
            if (j0 == qlen) { if (!(gscore > fc)) { max_ie = i; gscore = fc; } }   // sx:1768,1913,1941
            done = true; break;                                                    // sx:1942
        }
        if ((lim & ~(GROUP - 1)) + GROUP - (j0 & ~(GROUP - 1)) > rcap) { overflow = true; done = true; break; }

        int carry = 0;            // f entering the first column of the round
        int hcarry = fc;          // h of the column left of the round
        int key = -1;
        for (int gbase = j0 & ~(GROUP - 1); gbase <= lim; gbase += GROUP) {
            const int jl = gbase + 8 * sl;
            const int lo = j0 - jl, hi = lim - jl;              // columns k with lo <= k < hi are cells of this row
            const bool full = (lo <= 0) && (hi >= 8);
            const uint4 wa = *reinterpret_cast<const uint4*>(eh + (jl & rm));
            const uint4 wb = *reinterpret_cast<const uint4*>(eh + (jl & rm) + 4);
            const uint32_t wd[8] = { wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w };
            const uint32_t qw = (jl >> 3) < nqw ? __ldg(qs + (jl >> 3)) : 0u;
            const uint32_t x = GENERIC ? qw : (qw ^ trep);
            int hh[8], fl[8];
            int run = 0;
            // pass 1: everything that does not need the incoming F
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int M = (int)(wd[k] & 0xffffu), e = (int)(wd[k] >> 16);
                const int s = k2s_score<GENERIC>(GENERIC ? ((x >> (4 * k)) & 15u) : (x & (0xfu << (4 * k))), mat, mis, rlo, rhi);
                hh[k] = add_max(M, s, e);                                          // sx:1797,1798
                int g = add_max_relu(hh[k], -oe_ins, 0);                           // sx:1863,1865 with h >= hh
                if (!full && !(k >= lo && k < hi)) g = 0;
                fl[k] = run;
                run = add_max(run, -e_ins, g);                                     // sx:1780,1781
            }
            // carries across the sub-warp's lanes: prefix max of A[l] + 8*e_ins*l
            int P = run + e8 * sl;
#pragma unroll
            for (int d = 1; d < SW; d <<= 1) {
                const int o = __shfl_up_sync(smask, P, d, SW);
                if (sl >= d) P = imax(P, o);
            }
            const int Pex = __shfl_up_sync(smask, P, 1, SW);
            int fin = carry - e8 * sl;
            if (sl > 0) fin = imax(fin, Pex - e8 * (sl - 1));
            fin = imax(fin, 0);
            const int fout = imax(run, fin - e8);
            carry = __shfl_sync(smask, fout, SW - 1, SW);
            // pass 2: fold the carry, finish H, E, key, zero bits
            int h[8];
            uint32_t enew[8];
            uint32_t zbits = 0;
            int u = fin;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int f = imax(fl[k], u);
                u -= e_ins;
                h[k] = imax(hh[k], f);                                             // sx:1809
                const int t = add_max_relu(h[k], -oe_del, 0);                      // sx:1866,1862
                enew[k] = add_max_s16x2(wd[k], ce_pack, (uint32_t)t << 16);        // {max(e-e_del,t), 0}  sx:1770-1771
                if (full || (k >= lo && k < hi)) {
                    key = imax(key, h[k] * 65536 + jl + k);                        // sx:1808,1816
                    zbits |= (h[k] == 0 ? 1u : 0u) << k;
                }
            }
            int hleft = __shfl_up_sync(smask, h[7], 1, SW);
            if (sl == 0) hleft = hcarry;
            hcarry = __shfl_sync(smask, h[7], SW - 1, SW);
            if (full) {
                uint4 oa, ob;
                oa.x = enew[0] | (uint32_t)((lo == 0) ? fc : hleft);
                oa.y = enew[1] | (uint32_t)h[0]; oa.z = enew[2] | (uint32_t)h[1]; oa.w = enew[3] | (uint32_t)h[2];
                ob.x = enew[4] | (uint32_t)h[3]; ob.y = enew[5] | (uint32_t)h[4]; ob.z = enew[6] | (uint32_t)h[5]; ob.w = enew[7] | (uint32_t)h[6];
                *reinterpret_cast<uint4*>(eh + (jl & rm)) = oa;
                *reinterpret_cast<uint4*>(eh + (jl & rm) + 4) = ob;
            } else {
                // boundary lane: store cells [lo,hi) and the end slot eh[lim] = {h1, 0} (sx:1775,1904,1776)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (k >= lo && k <= hi) {
                        const int h1 = (k == lo) ? fc : (k ? h[k - 1] : hleft);
                        eh[(jl + k) & rm] = (k < hi ? enew[k] : 0u) | (uint32_t)h1;
                    }
                }
            }
            reinterpret_cast<unsigned char*>(zb)[(jl & rm) >> 3] = (unsigned char)zbits;
        }
        __syncwarp(smask);

        // ---- row epilogue (identical in every lane of the sub-warp) ----
        key = __reduce_max_sync(smask, key);
        const int m = key >> 16, mj = key & 0xffff;
        cells += (unsigned long long)(lim - j0);
        const int h1 = (int)(eh[lim & rm] & 0xffffu);
        if (lim == qlen) {                                                         // sx:1768,1913
            if (!(gscore > h1)) { max_ie = i; gscore = h1; }                       // sx:1941,1829,1831
        }
        if (m == 0) { done = true; break; }                                        // sx:1942
        if (m > max) {                                                             // sx:1959
            max = m; max_i = i; max_j = mj;
            const int d = mj > i ? mj - i : i - mj;
            max_off = max_off > d ? max_off : d;                                   // sx:1707-1708,1812
        } else if (zdrop > 0) {                                                    // ksw_extend2 z-drop (not in the RTL)
            const int di = i - max_i, dj = mj - max_j;
            if (di > dj) { if (max - m - (di - dj) * e_del > zdrop) { done = true; break; } }
            else         { if (max - m - (dj - di) * e_ins > zdrop) { done = true; break; } }
        }
        // narrowing (sx:1766-1769 / 1779,1782-1789): zero bit of cell c <-> eh[c+1].h == 0
        int cb = -1, ce = 0x7fffffff;
        for (int wbase = (j0 >> 5); wbase <= ((lim - 1) >> 5); wbase += SW) {
            const int wi = wbase + sl;
            const uint32_t zw = (wi <= ((lim - 1) >> 5)) ? zb[wi & (rm >> 5)] : 0u;
            const uint32_t za = zw & k2s_range_mask(wi * 32, j0, mj - 1);
            const uint32_t ze = zw & k2s_range_mask(wi * 32, mj + 1, lim - 1);
            if (za) cb = imax(cb, wi * 32 + 31 - __clz(za));
            if (ze) ce = imin(ce, wi * 32 + __ffs(ze) - 1);
        }
        cb = __reduce_max_sync(smask, cb);
        ce = __reduce_min_sync(smask, ce);
        beg = cb >= 0 ? cb + 2 : (fc == 0 ? j0 + 1 : j0);
        end = ce != 0x7fffffff ? ce + 1 : lim + 1;
      } while (0);
      __syncwarp();                                    // re-converge the four sub-warps (also orders the row buffer writes)
    }

    if (sl == 0 && has_task) {
        int4* o = reinterpret_cast<int4*>(A.out + (A.out_index ? A.out_index[slot] : slot));
        const unsigned long long cc = cells > 0x7fffffffull ? 0x7fffffffull : cells;
        o[0] = make_int4(max, max_j + 1, max_i + 1, max_ie + 1);                   // sx:1315-1375 (score,qle,tle,gtle)
        o[1] = make_int4(gscore, max_off, overflow ? 0 : (int)cc, overflow ? STATUS_OVERFLOW : STATUS_OK);
        if (A.cells_total && cells && !overflow) atomicAdd(A.cells_total, cells);
    }
}

// Ring size for a K2S launch: the first row (min(qmax, wmax+1) cells + end slot, 64-aligned) must fit; 0 = not applicable.
int k2s_ring_cols(int qmax, int wmax)
{
    const int first = ((qmax < wmax + 1 ? qmax : wmax + 1) + 1 + 63) & ~63;
    int r = 512;
    while (r < first) r <<= 1;
    return r <= K2S_RING_MAX ? r : 0;
}

template <int GENERIC>
static cudaError_t k2s_launch_t(const LaunchArgs& a, cudaStream_t st)
{
    if (!a.ntiles) return cudaSuccess;
    constexpr int SW = 8, TPW = 32 / SW;
    const int rcap = a.ring_cols;
    if (rcap < 64 || (rcap & (rcap - 1))) return cudaErrorInvalidValue;
    const size_t smem = (size_t)K2S_HDR_BYTES + (size_t)TPW * k2s_task_words(rcap) * 4u + 16u;
    auto kern = k2s_extend_kernel<GENERIC, SW>;
    if (smem > 232448) return cudaErrorInvalidValue;
    static std::atomic<unsigned> smem_set{ 0u };          // per instantiation of this launcher
    cudaError_t err = ensure_max_smem(kern, smem_set);
    if (err != cudaSuccess) return err;
    kern<<<(a.ntiles + TPW - 1) / TPW, 32, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t k2s_launch(const LaunchArgs& a, int generic, cudaStream_t st)
{
    return generic ? k2s_launch_t<1>(a, st) : k2s_launch_t<0>(a, st);
}

}  // namespace bsw
