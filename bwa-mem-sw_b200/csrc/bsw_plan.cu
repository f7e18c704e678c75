// bsw_plan.cu -- the task scheduler on the device (sm_100a): sort a chunk's tasks by shape and lay out K1 tiles.
//
// sw_pe_array_task_parse (sw_pe_array_task_parse.v:1600-1650,1652-1762) hands every task to the next PE with room; the
// B200 counterpart groups the 32 tasks of a warp by shape.  Round 1 did that on the host (radix sort + tile building,
// ~18 ns per task per host thread, bsw_sched.cpp::build_plan) and shipped tile / slot / index arrays; with eight GPUs
// on one box the host became the bottleneck.  Here the host uploads the chunk's tasks in input order and only keeps
// per-bucket counts (enough to size the launches); the device does the rest:
// Three small, fully parallel kernels per chunk (a counting sort on the chunk's 20-bit key -- the same key as
// build_plan's chunk key: [matrix class:1][qlen/16:7][min(qlen,h0)/4:6][tlen/8:6], descending cost):
//   dp_count    one thread per task: key -> (major bucket, 12-bit sub key); count the bin
//   dp_scan     one CTA per NON-EMPTY major bucket (class, qlen/16): exclusive scan of its 4096 bins, offset by the
//               bucket's first sorted position, which the host knows from its per-bucket counts
//   dp_place    one thread per task: position = atomicAdd(bin); writes the slot scalars, the source offsets and the
//               slot -> task index (padding lanes zeroed).  Tasks of one bin land in arbitrary order: which tasks share
//               a warp never changes a result (tests: permutation invariance), only equal-cost tasks swap places
// The K0 gather, which reads every slot of a tile anyway, then takes the tile's word counts (warp max) into the tile
// header the host pre-filled with offsets.  Nothing comes back to the host.
// Two earlier versions, kept here as a warning: a stable LSD radix sort with one kernel per step (7 dependent,
// latency-bound launches in front of every chunk: batch call 5.9 -> 6.7 ms per 1 M tasks although the host time fell),
// and the same sort as ONE 1024-thread CTA holding the keys in registers (62 registers x 1024 threads = the whole
// register file of an SM: the CTA starves behind the K1 CTAs that keep refilling every SM, again 6.8 ms).
#include <cuda_runtime.h>
#include "bsw_device.cuh"
#include "bsw_kernels.h"

namespace bsw {

constexpr int DP_SUB = 4096;                 // bins per major bucket: [63 - min(qlen,h0)/4 : 6][63 - tlen/8 : 6]

__device__ __forceinline__ uint32_t dp_bin(const DpArgs& A, const SlotParam& t, uint32_t cls)
{
    const uint32_t qb = (uint32_t)min(t.qlen >> 4, 127);
    const uint32_t wd = (uint32_t)min(min(t.qlen, t.h0) >> 2, 63), tl = (uint32_t)min(t.tlen >> 3, 63);
    return (uint32_t)A.major_of[((cls & 1u) << 7) | qb] * DP_SUB + (((63u - wd) << 6) | (63u - tl));
}

__global__ void __launch_bounds__(256) dp_count_kernel(const __grid_constant__ DpArgs A)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= A.count) return;
    const uint32_t b = dp_bin(A, A.task_param[i], A.task_cls ? A.task_cls[i] : A.const_cls);
    A.task_bin[i] = b;
    atomicAdd(A.bins + b, 1u);
}

__global__ void __launch_bounds__(256) dp_scan_kernel(const __grid_constant__ DpArgs A)
{
    __shared__ uint32_t wsum[8];
    uint32_t* bins = A.bins + (size_t)blockIdx.x * DP_SUB;
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    uint32_t v[16], sum = 0;
    const uint4* src = reinterpret_cast<const uint4*>(bins + tid * 16);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const uint4 q = src[k]; v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w; sum += q.x + q.y + q.z + q.w; }
    uint32_t incl = sum;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[wp] = incl;
    __syncthreads();
    uint32_t acc = A.major_start[blockIdx.x] + incl - sum;
    for (int w = 0; w < wp; ++w) acc += wsum[w];
    uint4* dst = reinterpret_cast<uint4*>(bins + tid * 16);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint4 q;
        q.x = acc; acc += v[4 * k]; q.y = acc; acc += v[4 * k + 1]; q.z = acc; acc += v[4 * k + 2]; q.w = acc; acc += v[4 * k + 3];
        dst[k] = q;
    }
}

__global__ void __launch_bounds__(256) dp_place_kernel(const __grid_constant__ DpArgs A)
{
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (blockIdx.x == 0 && threadIdx.x < 64) {                         // padding lanes of each class's last tile
        const int c = threadIdx.x >> 5;
        const uint32_t slot = A.class_slot0[c] + A.class_count[c] + (threadIdx.x & 31u);
        const uint32_t end = A.class_slot0[c] + ((A.class_count[c] + TILE_LANES - 1) / TILE_LANES) * TILE_LANES;
        if (slot < end) { A.slots[slot] = SlotParam{ 0, 0, 0, 0 }; A.slot_src[slot] = SlotSrc{ 0, 0 }; A.out_index[slot] = 0xffffffffu; }
    }
    if (i >= A.count) return;
    const uint32_t p = atomicAdd(A.bins + A.task_bin[i], 1u);          // sorted position (class-major, descending cost)
    const int c = (p >= A.class_pos0[1]) ? 1 : 0;
    const uint32_t slot = A.class_slot0[c] + (p - A.class_pos0[c]);
    A.slots[slot] = A.task_param[i];
    A.slot_src[slot] = A.task_src[i];
    A.out_index[slot] = i;
}

cudaError_t dp_plan_launch(const DpArgs& a, cudaStream_t st)
{
    if (!a.count) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(a.bins, 0, (size_t)a.nmajor * DP_SUB * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    const uint32_t nblk = (a.count + 255) / 256;
    dp_count_kernel<<<nblk, 256, 0, st>>>(a);
    dp_scan_kernel<<<a.nmajor, 256, 0, st>>>(a);
    dp_place_kernel<<<nblk, 256, 0, st>>>(a);
    return cudaGetLastError();
}

size_t dp_bins_per_major() { return DP_SUB; }

}  // namespace bsw
