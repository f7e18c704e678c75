// bsw_k0.cu -- K0: tile gather, the device half of the task scheduler (sm_100a).
//
// The host packs every sequence exactly once, in input order, into a task-major source arena (sequential reads and
// writes at memory speed, bsw_sched.cpp::pack_tasks) and only sorts task *indices*.  This kernel plays the part of
// sw_pe_array_task_parse's data path (sw_pe_array_task_parse.v:924-948: fetch each task's words from the batch buffer
// and hand them to its PE): one warp per K1 tile copies the 32 tasks' packed words into the tile-interleaved layout
// q[k*32+lane] / t[k*32+lane] that K1 consumes with one TMA bulk copy and coalesced 128-byte target loads.
// Reads are 128-bit per lane (each lane walks its own task), writes are full 128-byte lines.
#include <cuda_runtime.h>
#include "bsw_device.cuh"
#include "bsw_k1_core.cuh"
#include "bsw_kernels.h"

namespace bsw {

constexpr int K0_WARPS = 4;

__device__ __forceinline__ void k0_copy_block(const uint4* __restrict__ src16, int own_words, uint32_t* __restrict__ dst,
                                              int tile_words, int lane)
{
    // 32 bytes (one DRAM sector) per lane per step: a lane walks its own task, so a 16-byte step would fetch every
    // sector twice
    for (int m = 0; m * 4 < tile_words; m += 2) {
        uint4 v0 = make_uint4(0u, 0u, 0u, 0u), v1 = v0;
        if (m * 4 < own_words) v0 = __ldg(src16 + m);
        if (m * 4 + 4 < own_words) v1 = __ldg(src16 + m + 1);
        uint32_t* d = dst + (size_t)(m * 4) * TILE_LANES + lane;
        const uint32_t w[8] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w };
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (m * 4 + k < tile_words) d[k * TILE_LANES] = w[k];
    }
}

// Raw mode: the lane's sequence is `len` bytes (one base code each) at an arbitrary byte address.  Eight bases per step:
// one aligned 8-byte load (the previous one supplies the low part), nibble pack, and two SWAR tests -- any byte above 4,
// any byte equal to 4.  Reads run at most 15 bytes past the sequence (the raw buffers carry that slack).
__device__ __forceinline__ uint32_t k0_pack_raw(const uint8_t* __restrict__ base, int len, uint32_t* __restrict__ dst,
                                                int tile_words, int lane)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(base);
    const unsigned long long* p = reinterpret_cast<const unsigned long long*>(a & ~(uintptr_t)7);
    const int sh = (int)(a & 7u) * 8;
    uint32_t flags = 0;
    unsigned long long lo = len > 0 ? __ldg(p) : 0ull;
    for (int k = 0; k < tile_words; ++k) {
        uint32_t w = 0;
        if (8 * k < len) {
            const unsigned long long hi = __ldg(p + k + 1);
            unsigned long long v = sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
            lo = hi;
            const int rem = len - 8 * k;
            if (rem < 8) v &= (1ull << (8 * rem)) - 1ull;
            if ((v | (v + 0x7b7b7b7b7b7b7b7bull)) & 0x8080808080808080ull) flags |= SLOT_BAD_CODE;       // some byte >= 5
            const unsigned long long z = v ^ 0x0404040404040404ull;                                      // zero byte <=> code 4
            if ((z - 0x0101010101010101ull) & ~z & 0x8080808080808080ull) flags |= SLOT_HAS_N;
            unsigned long long x = (v | (v >> 4)) & 0x00ff00ff00ff00ffull;
            x = (x | (x >> 8)) & 0x0000ffff0000ffffull;
            x = (x | (x >> 16));
            w = (uint32_t)x;
        }
        dst[(size_t)k * TILE_LANES + lane] = w;
    }
    return flags;
}

// 2 bit per base source (bsw_pack2.cpp): 16 bases per word, sequences start on a word.  Every source word becomes two
// tile words (8 bases each, 4 bit per base): the 2-bit fields are spread to nibbles with three shift-and-mask steps.
__device__ __forceinline__ uint32_t k0_spread16(uint32_t x)
{
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    return x;
}
__device__ __forceinline__ void k0_unpack2(const uint32_t* __restrict__ src, int len, uint32_t* __restrict__ dst, int tile_words, int lane)
{
    for (int m = 0; 2 * m < tile_words; ++m) {
        const uint32_t w = (16 * m < len) ? __ldg(src + m) : 0u;
        dst[(size_t)(2 * m) * TILE_LANES + lane] = k0_spread16(w & 0xffffu);
        if (2 * m + 1 < tile_words) dst[(size_t)(2 * m + 1) * TILE_LANES + lane] = k0_spread16(w >> 16);
    }
}

__global__ void __launch_bounds__(K0_WARPS * 32) k0_gather_kernel(const __grid_constant__ GatherArgs A)
{
    const int lane = threadIdx.x & 31;
    const uint32_t tile = blockIdx.x * K0_WARPS + (threadIdx.x >> 5);
    if (tile >= A.ntiles) return;
    const TileHdr hd = A.tiles[tile];
    int nqw = (int)(hd.nqw_ntw & 0x7fffu), ntw = (int)(hd.nqw_ntw >> 16);
    const SlotParam sp = A.slots[hd.slot0 + lane];
    if (A.dp_tiles) {
        // the device planner placed the tasks; the header's counts are the host's upper bounds (they size the arena):
        // take the real ones so that K1 moves and walks no more words than the tile has
        int tq = sp.qlen, tt = sp.tlen;
        for (int o = 16; o; o >>= 1) { tq = max(tq, __shfl_xor_sync(0xffffffffu, tq, o)); tt = max(tt, __shfl_xor_sync(0xffffffffu, tt, o)); }
        nqw = (tq + 7) >> 3; ntw = (tt + 7) >> 3;
        if (lane == 0) A.dp_tiles[tile].nqw_ntw = (uint32_t)nqw | ((uint32_t)ntw << 16);
    }
    const SlotSrc ss = A.slot_src[hd.slot0 + lane];
    const int own_q = sp.qlen > 0 ? ((sp.qlen + 31) >> 5) * 4 : 0;      // words the host packed for this task (zero padded)
    const int own_t = sp.tlen > 0 ? ((sp.tlen + 31) >> 5) * 4 : 0;
    if (A.src2) {
        k0_unpack2(A.src2 + ss.qoff16, sp.qlen, A.dst + (size_t)hd.qoff16 * 4, nqw, lane);
        k0_unpack2(A.src2 + ss.toff16, sp.tlen, A.dst + (size_t)hd.toff16 * 4, ntw, lane);
        return;
    }
    if (A.raw_q) {
        uint32_t f = k0_pack_raw(A.raw_q + ss.qoff16, sp.qlen, A.dst + (size_t)hd.qoff16 * 4, nqw, lane);
        f |= k0_pack_raw(A.raw_t + ss.toff16, sp.tlen, A.dst + (size_t)hd.toff16 * 4, ntw, lane);
        A.slot_flags[hd.slot0 + lane] = f;
        return;
    }
    const uint4* src16 = reinterpret_cast<const uint4*>(A.src);
    k0_copy_block(src16 + ss.qoff16, own_q, A.dst + (size_t)hd.qoff16 * 4, nqw, lane);
    k0_copy_block(src16 + ss.toff16, own_t, A.dst + (size_t)hd.toff16 * 4, ntw, lane);
}

cudaError_t k0_launch(const GatherArgs& a, cudaStream_t st)
{
    if (!a.ntiles) return cudaSuccess;
    k0_gather_kernel<<<(a.ntiles + K0_WARPS - 1) / K0_WARPS, K0_WARPS * 32, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace bsw
