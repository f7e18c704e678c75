// bsw_kernels.h -- host-callable launchers of the extension kernels.
#pragma once
#include <cuda_runtime.h>
#include "bsw_device.cuh"
#include <atomic>

namespace bsw {
// Opt the kernel into the full 227 KB of dynamic shared memory, once per kernel instantiation and device.  Always the
// same (maximal) value: launches come concurrently from several host threads, a per-launch value would race; and a
// driver call per launch is what the host workers contend on (22 -> ~10 API calls per chunk with this and the
// narrower stream fork).
// `done` is the calling launcher's own static (one per kernel instantiation): one bit per device id < 32.
template <class Kernel>
inline cudaError_t ensure_max_smem(Kernel kern, std::atomic<unsigned>& done)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned bit = 1u << (dev & 31);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}
}  // namespace bsw


namespace bsw {

// K0: tile gather (source arena -> tiled arena) for the K1 tiles of a chunk.
cudaError_t k0_launch(const GatherArgs& a, cudaStream_t st);

// Device-side scheduler: keys, radix sort, tile building for the K1 tasks of a chunk (bsw_plan.cu).
cudaError_t dp_plan_launch(const DpArgs& a, cudaStream_t st);
size_t dp_bins_per_major();     // u32 counters per non-empty (class, qlen/16) bucket in DpArgs.bins

// K1: inter-task kernel (one lane per task, one warp-tile per CTA).  variant 1|2 (recurrence policy),
// generic 0|1 (5x5 matrix lookup instead of match/mismatch), sym 0|1 (o_del==o_ins && e_del==e_ins).
cudaError_t k1_launch(const LaunchArgs& a, int variant, int generic, int sym, cudaStream_t st);
size_t k1_smem_bytes(int qmax, int nqw_max);

// K3: fused seed-task kernel (level 2 on the device): left + right extension, band retry, clip, record.
// a.tiles = (left, right) tile pairs, a.seeds[pair*32+lane], a.out[pair*32+lane] = bsw_aln_record.
cudaError_t k3_launch(const LaunchArgs& a, int variant, int generic, int sym, cudaStream_t st);
size_t k3_smem_bytes(int qmax, int nqw_max);

// K2: intra-task kernel (one warp per task, row-parallel with a prefix-max scan for F).  variant 1|2.
cudaError_t k2_launch(const LaunchArgs& a, int generic, int warps, int variant, cudaStream_t st);
size_t k2_smem_bytes(int qmax, int wmax);


// K4: banded global alignment with traceback (ksw_global2), one lane per task.
cudaError_t k4_launch(const GlobalArgs& a, cudaStream_t st);

// K5: extension with 32-bit row state in global memory, one warp per task (tasks outside the 16-bit envelope of K1 / K2).
cudaError_t k5_launch(const WideArgs& a, int variant, int sm_count, cudaStream_t st);

// INT-pipe micro-benchmark (roofline denominator): runs `iters` rounds of dependent-free instruction
// streams on every SM; out_ops[5] = {add, max, fused add-max (x2 ops), DP-cell mix, add on both pipes} in ops per second.
cudaError_t int_peak_run(double out_ops[5], double* sm_clock_mhz, int* sm_count, cudaStream_t st);

}  // namespace bsw
