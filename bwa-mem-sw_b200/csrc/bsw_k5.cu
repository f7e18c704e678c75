// bsw_k5.cu -- K5: the wide extension kernel, one warp per task, 32-bit row state in global memory (sm_100a).
//
// K1 and K2 keep H / E in 16 bits and are exact while h0 + qlen * max(mat) <= 32767 and qlen <= 40000
// (bsw_sched.h: SCORE_CAP, K2_QLEN_CAP).  Tasks outside that envelope (reads of tens of kilobases whose seed chain has
// already scored more than 32767, or queries longer than K2's shared-memory row) come here.  Same recurrence as
// ksw_extend2 / sw_extend (sw_pe_array_sw_extend.v:1891-1981; V1 and V2 as in bsw_k1_core.cuh), same row-buffer slot
// semantics (slot j holds H(i-1, j-1) and E(i, j); slot `end` is written after the row), so the band narrowing -- which
// reads slots the current row did not touch -- sees what the scalar code sees.
//
// A row is processed 32 columns at a time.  The F chain f(j+1) = max(f(j) - e_ins, g(j)) with g(j) = max(x(j) - oe_ins, 0),
// x = the cell before F is applied (V1; o_ins >= 0 makes the F term of h redundant) or M (V2), is a max-plus prefix scan:
// f(j+1) = max_{k<=j} (g(k) + k*e_ins) - j*e_ins, five shuffles per 32 columns plus a carry between the chunks.
// Rare by construction, so it is written for exactness, not for the ALU roofline: about 4 coalesced 128-byte accesses
// per 32 cells against L2-resident rows.
#include <cuda_runtime.h>
#include "bsw_device.cuh"
#include "bsw_kernels.h"

namespace bsw {

namespace {

constexpr int K5_WARPS = 4;
constexpr int K5_NEG = -0x3fffffff;

__device__ __forceinline__ int k5_warp_max(int v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}

template <int VARIANT>
__device__ void k5_task(const WideArgs& A, const WideTask& g, int lane, SlotResult* res)
{
    const uint8_t* __restrict__ query = A.qbuf + g.qoff;
    const uint8_t* __restrict__ target = A.tbuf + g.toff;
    int32_t* H = A.rows + g.row_off;            // qlen + 2 slots
    int32_t* E = H + (g.qlen + 2);
    const int qlen = g.qlen, tlen = g.tlen, w = g.w, h0 = g.h0;
    const int o_del = A.p.o_del, e_del = A.p.e_del, e_ins = A.p.e_ins;
    const int oe_del = o_del + e_del, oe_ins = A.p.o_ins + e_ins, zdrop = A.p.zdrop;

    // first row: slot j = H(-1, j-1), all E = 0
    for (int j = lane; j <= qlen + 1; j += 32) {
        int v = 0;
        if (j == 0) v = h0;
        else if (j <= qlen) { const long long t = (long long)h0 - oe_ins - (long long)(j - 1) * e_ins; v = t > 0 ? (int)t : 0; }
        H[j] = v; E[j] = 0;
    }
    __syncwarp();

    int max_sc = h0, max_i = -1, max_j = -1, max_ie = -1, gscore = -1, max_off = 0;
    int beg = 0, end = qlen;
    unsigned long long ncell = 0;
    for (int i = 0; i < tlen; ++i) {
        const int8_t* srow = A.p.mat + 5 * target[i];
        if (beg < i - w) beg = i - w;
        if (end > i + w + 1) end = i + w + 1;
        if (end > qlen) end = qlen;
        int h1;                                                   // first column
        if (VARIANT == 1 || beg == 0) { const long long t = (long long)h0 - (o_del + (long long)e_del * (i + 1)); h1 = t > 0 ? (int)t : 0; }
        else h1 = 0;
        int m = 0, mj = -1;
        int hcarry = h1;                                          // h of the column left of the chunk
        int fcarry = 0;                                           // f entering the chunk's first column
        for (int c = beg; c < end; c += 32) {
            const int j = c + lane;
            const bool on = j < end;
            int M = 0, e = 0, sc = 0;
            if (on) { M = H[j]; e = E[j]; sc = srow[query[j]]; }
            if (VARIANT == 1) M = M + sc; else M = M ? M + sc : 0;
            const int x = max(M, e);                              // the cell without its F term
            int gk = (VARIANT == 1 ? x : M) - oe_ins;
            gk = gk > 0 ? gk : 0;
            int u = on ? gk + j * e_ins : K5_NEG;                 // inclusive max scan of g(k) + k * e_ins
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(0xffffffffu, u, d); if (lane >= d) u = max(u, o); }
            int ux = __shfl_up_sync(0xffffffffu, u, 1);           // exclusive: columns c .. j-1
            int f = fcarry - lane * e_ins;
            if (lane > 0) f = max(f, ux - (j - 1) * e_ins);
            const int h = max(x, f);
            // next chunk's carries
            const int u31 = __shfl_sync(0xffffffffu, u, 31);
            fcarry = max(fcarry - 32 * e_ins, u31 - (c + 31) * e_ins);
            // E(i+1, j)
            int t = (VARIANT == 1 ? h : M) - oe_del;
            t = t > 0 ? t : 0;
            e = max(e - e_del, t);
            // slot j takes H(i, j-1)
            int hl = __shfl_up_sync(0xffffffffu, h, 1);
            if (lane == 0) hl = hcarry;
            if (on) { H[j] = hl; E[j] = e; }                      // a lane reads and writes only its own slot: no hazard inside the row
            const int last = min(31, end - 1 - c);
            hcarry = __shfl_sync(0xffffffffu, h, last);
            // running maximum: the LAST column that reaches it (ties move mj right)
            const int hm = on ? h : -1;
            const int cm = k5_warp_max(hm);
            if (cm >= m) {
                const unsigned hit = __ballot_sync(0xffffffffu, on && h == cm);
                m = cm; mj = c + 31 - __clz(hit);
            }
        }
        if (lane == 0) { H[end] = hcarry; E[end] = 0; }
        __syncwarp();
        if (end > beg) ncell += (unsigned long long)(end - beg);
        const int jfin = end > beg ? end : beg;
        if (jfin == qlen) {
            max_ie = gscore > hcarry ? max_ie : i;
            gscore = gscore > hcarry ? gscore : hcarry;
        }
        if (m == 0) break;
        if (m > max_sc) {
            max_sc = m; max_i = i; max_j = mj;
            const int d = mj > i ? mj - i : i - mj;
            max_off = max_off > d ? max_off : d;
        } else if (zdrop > 0) {
            const int di = i - max_i, dj = mj - max_j;
            if (di > dj) { if (max_sc - m - (di - dj) * e_del > zdrop) break; }
            else { if (max_sc - m - (dj - di) * e_ins > zdrop) break; }
        }
        // narrowing for the next row, on the row buffer as it now stands
        if (VARIANT == 1) {
            int nb = beg;                                         // first zero at or left of mj
            for (int c = mj; c >= beg; c -= 32) {
                const int j = c - lane;
                const unsigned z = __ballot_sync(0xffffffffu, j >= beg && H[j] == 0);
                if (z) { nb = c - (__ffs(z) - 1) + 1; break; }
            }
            int ne = end + 1;                                     // first zero at or right of mj + 2 (the scan stops past `end`)
            for (int c = mj + 2; c <= end; c += 32) {
                const int j = c + lane;
                const unsigned z = __ballot_sync(0xffffffffu, j <= end && H[j] == 0);
                if (z) { ne = c + __ffs(z) - 1; break; }
            }
            if (mj + 2 > end) ne = mj + 2;
            beg = nb; end = ne;
        } else {
            int nb = end;
            for (int c = beg; c < end; c += 32) {
                const int j = c + lane;
                const unsigned nz = __ballot_sync(0xffffffffu, j < end && (H[j] != 0 || E[j] != 0));
                if (nz) { nb = c + __ffs(nz) - 1; break; }
            }
            int jl = nb - 1;                                      // last slot in [nb, end] that is not all zero
            for (int c = end; c >= nb; c -= 32) {
                const int j = c - lane;
                const unsigned nz = __ballot_sync(0xffffffffu, j >= nb && (H[j] != 0 || E[j] != 0));
                if (nz) { jl = c - (__ffs(nz) - 1); break; }
            }
            beg = nb;
            end = jl + 2 < qlen ? jl + 2 : qlen;
        }
    }
    if (lane == 0) {
        res->score = max_sc; res->qle = max_j + 1; res->tle = max_i + 1; res->gtle = max_ie + 1;
        res->gscore = gscore; res->max_off = max_off;
        res->cells = (int32_t)(ncell > 0xffffffffull ? 0xffffffffu : (uint32_t)ncell);
        res->status = STATUS_OK;
    }
}

template <int VARIANT>
__global__ void __launch_bounds__(K5_WARPS * 32) k5_wide_kernel(const __grid_constant__ WideArgs A)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * K5_WARPS;
    for (uint32_t task = blockIdx.x * K5_WARPS + (threadIdx.x >> 5); task < A.ntasks; task += warps)
        k5_task<VARIANT>(A, A.tasks[task], lane, A.out + task);
}

}  // namespace

cudaError_t k5_launch(const WideArgs& a, int variant, int sm_count, cudaStream_t st)
{
    if (!a.ntasks) return cudaSuccess;
    const uint32_t want = (a.ntasks + K5_WARPS - 1) / K5_WARPS;
    const uint32_t grid = want < (uint32_t)sm_count * 8u ? want : (uint32_t)sm_count * 8u;
    if (variant == 2) k5_wide_kernel<2><<<grid, K5_WARPS * 32, 0, st>>>(a);
    else k5_wide_kernel<1><<<grid, K5_WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace bsw
