// bsw_k4.cu -- K4: banded global alignment with traceback (ksw_global2), one lane per task (sm_100a).
//
// the DP that follows seed extension in BWA-MEM (bwa_gen_cigar2 -> ksw_global2, once per reported
// alignment).  It is not part of the reference tree; the recurrence and the traceback encoding follow the published BWA
// algorithm (see oracle/ksw_extend_ref.c::bswref_global, which carries the same note: parity unpinned for this row).
// Layout: a tile = 32 tasks, one per lane; row state {H, E} as two int32 per column and the direction bytes live in
// global memory, interleaved by lane (index * 32 + lane), so the lanes of a warp -- which walk their tasks in lockstep --
// touch one 128-byte line (ints) or one 32-byte sector (direction bytes) per step.  Bases are read one per byte from
// the caller's concatenated buffers.  The traceback runs on the same lane and emits run-length encoded operations in
// BAM encoding (len << 4 | op; 0 = M, 1 = I, 2 = D), reversed in place at the end.
#include <cuda_runtime.h>
#include "bsw_device.cuh"
#include "bsw_kernels.h"

namespace bsw {

constexpr int K4_MINUS_INF = -0x40000000;

__device__ __forceinline__ int k4_push(uint32_t* cigar, int n, int max_ops, int op, int len)
{
    if (n < 0) return n;
    if (n > 0 && (cigar[n - 1] & 0xfu) == (uint32_t)op) { cigar[n - 1] += (uint32_t)len << 4; return n; }
    if (n >= max_ops) return -1;
    cigar[n] = ((uint32_t)len << 4) | (uint32_t)op;
    return n + 1;
}

__global__ void __launch_bounds__(32) k4_global_kernel(const __grid_constant__ GlobalArgs A)
{
    const int lane = threadIdx.x;
    const uint32_t tile = blockIdx.x;
    const uint32_t task = tile * TILE_LANES + lane;
    if (task >= A.ntasks) return;
    const GlobalTask g = A.tasks[task];
    const int qlen = g.qlen, tlen = g.tlen, w = g.w;
    const uint8_t* query = A.qbuf + g.qoff;
    const uint8_t* target = A.tbuf + g.toff;
    int32_t* eh = A.eh + (size_t)A.eh_off[tile] * TILE_LANES + lane;          // eh[(2*j + {0: h, 1: e}) * 32]
    uint8_t* z = A.z + (size_t)A.z_off[tile] * TILE_LANES + lane;            // z[(i * n_col + c) * 32]
    const int o_del = A.p.o_del, e_del = A.p.e_del, o_ins = A.p.o_ins, e_ins = A.p.e_ins;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
#define K4_H(J) eh[(size_t)(2 * (J)) * TILE_LANES]
#define K4_E(J) eh[(size_t)(2 * (J) + 1) * TILE_LANES]
    K4_H(0) = 0; K4_E(0) = K4_MINUS_INF;
    int j = 1;
    for (; j <= qlen && j <= w; ++j) { K4_H(j) = -(o_ins + e_ins * j); K4_E(j) = K4_MINUS_INF; }
    for (; j <= qlen; ++j) { K4_H(j) = K4_MINUS_INF; K4_E(j) = K4_MINUS_INF; }
    for (int i = 0; i < tlen; ++i) {
        int f = K4_MINUS_INF;
        const int8_t* srow = A.p.mat + 5 * target[i];
        const int beg = i > w ? i - w : 0;
        const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
        int h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : K4_MINUS_INF;
        uint8_t* zi = z + (size_t)i * n_col * TILE_LANES;
        for (j = beg; j < end; ++j) {
            int m = K4_H(j), e = K4_E(j);
            K4_H(j) = h1;
            m += srow[query[j]];
            uint32_t d = m >= e ? 0u : 1u;
            int h = m >= e ? m : e;
            d = h >= f ? d : 2u;
            h = h >= f ? h : f;
            h1 = h;
            int t = m - oe_del;
            e -= e_del;
            d |= e > t ? 1u << 2 : 0u;
            e = e > t ? e : t;
            K4_E(j) = e;
            t = m - oe_ins;
            f -= e_ins;
            d |= f > t ? 2u << 4 : 0u;
            f = f > t ? f : t;
            zi[(size_t)(j - beg) * TILE_LANES] = (uint8_t)d;
        }
        K4_H(end) = h1; K4_E(end) = K4_MINUS_INF;
    }
    const int score = K4_H(qlen);
    uint32_t* cigar = A.cigar + (size_t)task * A.max_ops;
    int n = 0, which = 0;
    int i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
    while (i >= 0 && k >= 0 && n >= 0) {
        const uint32_t d = z[((size_t)i * n_col + (size_t)(k - (i > w ? i - w : 0))) * TILE_LANES];
        which = (int)((d >> (which << 1)) & 3u);
        if (which == 0) { n = k4_push(cigar, n, A.max_ops, 0, 1); --i; --k; }
        else if (which == 1) { n = k4_push(cigar, n, A.max_ops, 2, 1); --i; }
        else { n = k4_push(cigar, n, A.max_ops, 1, 1); --k; }
    }
    if (i >= 0) n = k4_push(cigar, n, A.max_ops, 2, i + 1);
    if (k >= 0) n = k4_push(cigar, n, A.max_ops, 1, k + 1);
    for (int a = 0; n > 0 && a < (n >> 1); ++a) { const uint32_t tmp = cigar[a]; cigar[a] = cigar[n - 1 - a]; cigar[n - 1 - a] = tmp; }
    A.score[task] = score;
    A.n_cigar[task] = n;
#undef K4_H
#undef K4_E
}

cudaError_t k4_launch(const GlobalArgs& a, cudaStream_t st)
{
    if (!a.ntasks) return cudaSuccess;
    k4_global_kernel<<<(a.ntasks + TILE_LANES - 1) / TILE_LANES, 32, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace bsw
