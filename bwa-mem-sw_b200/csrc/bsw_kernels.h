// bsw_kernels.h -- host-callable launchers of the extension kernels.
#pragma once
#include "bsw_device.cuh"

namespace bsw {

// K1: inter-task kernel (one thread per task).  variant 1|2, generic 0|1 (scoring), sym 0|1
// (o_del==o_ins && e_del==e_ins), wide 0|1 (int32 row buffer).
cudaError_t k1_launch(const LaunchArgs& a, int variant, int generic, int sym, int wide, cudaStream_t st);
size_t k1_smem_bytes(int qmax, int wide);

// K2: intra-task kernel (one warp per task, row-parallel with a prefix-max scan for F).
cudaError_t k2_launch(const LaunchArgs& a, int variant, int generic, cudaStream_t st);
size_t k2_smem_bytes(int qmax);

}  // namespace bsw
