// bsw_k1p.cu -- K1P: inter-task extension kernel, TWO extension tasks per lane with int16x2-packed scores (sm_100a).
//
// One CTA = one warp = one pair of K1 tiles (A tile, B tile; lane l carries tasks 2l and 2l+1 of the sorted order).
// Both packed query blocks arrive by TMA bulk copies (cp.async.bulk -> UBLKCP) on one mbarrier; the targets stream with
// coalesced 128-byte loads.  The DP is in bsw_k1p_core.cuh; it is bit-identical to K1 (bsw_k1_core.cuh) task by task.
#include <cuda_runtime.h>
#include "bsw_device.cuh"
#include "bsw_k1p_core.cuh"
#include "bsw_kernels.h"

namespace bsw {

constexpr int K1P_NT = TILE_LANES;
constexpr int K1P_HDR_BYTES = 128;

__device__ __forceinline__ uint32_t k1p_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int SYM>
__global__ void __launch_bounds__(K1P_NT) k1p_extend_kernel(const __grid_constant__ LaunchArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const TileHdr ha = A.tiles[2 * blockIdx.x], hb = A.tiles[2 * blockIdx.x + 1];
    const uint32_t nqa = ha.nqw_ntw & 0x7fffu, nqb = hb.nqw_ntw & 0x7fffu;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    const size_t qwords = (size_t)(A.nqw_max + K1_QS_EXTRA) * K1P_NT;
    uint32_t* qsa = reinterpret_cast<uint32_t*>(smem_raw + K1P_HDR_BYTES);
    uint32_t* qsb = qsa + qwords;
    uint32_t* eh = qsb + qwords;                                   // uint2 per column per lane
    const uint32_t bytes_a = nqa * K1P_NT * 4u, bytes_b = nqb * K1P_NT * 4u;

    if (lane == 0) {
        const uint32_t bar = k1p_smem_u32(mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes_a + bytes_b) : "memory");
        if (bytes_a) {
            const void* src = reinterpret_cast<const uint4*>(A.arena) + ha.qoff16;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(k1p_smem_u32(qsa)), "l"(src), "r"(bytes_a), "r"(bar) : "memory");
        }
        if (bytes_b) {
            const void* src = reinterpret_cast<const uint4*>(A.arena) + hb.qoff16;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(k1p_smem_u32(qsb)), "l"(src), "r"(bytes_b), "r"(bar) : "memory");
        }
    }
    const uint32_t slot_a = ha.slot0 + lane, slot_b = hb.slot0 + lane;
    const SlotParam spa = A.slots[slot_a], spb = A.slots[slot_b];
    __syncwarp();
    {
        const uint32_t bar = k1p_smem_u32(mbar);
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar) : "memory");
        }
    }

    unsigned long long my_cells = 0;
    if (spa.qlen > 0 || spb.qlen > 0) {
        const uint32_t* tga = A.arena + (size_t)ha.toff16 * 4u + lane;
        const uint32_t* tgb = A.arena + (size_t)hb.toff16 * 4u + lane;
        SlotResult ra, rb;
        k1p_pair<SYM>(A.p, spa, spb, (int)nqa, (int)nqb, eh + 2 * lane, qsa + lane, qsb + lane, tga, tgb, ra, rb);
        if (spa.qlen > 0) {
            int4* o = reinterpret_cast<int4*>(A.out + (A.out_index ? A.out_index[slot_a] : slot_a));
            o[0] = make_int4(ra.score, ra.qle, ra.tle, ra.gtle);
            o[1] = make_int4(ra.gscore, ra.max_off, ra.cells, ra.status);
            my_cells += (uint32_t)ra.cells;
        }
        if (spb.qlen > 0) {
            int4* o = reinterpret_cast<int4*>(A.out + (A.out_index ? A.out_index[slot_b] : slot_b));
            o[0] = make_int4(rb.score, rb.qle, rb.tle, rb.gtle);
            o[1] = make_int4(rb.gscore, rb.max_off, rb.cells, rb.status);
            my_cells += (uint32_t)rb.cells;
        }
    }
    if (A.cells_total) {
        for (int o = 16; o; o >>= 1) my_cells += __shfl_xor_sync(0xffffffffu, my_cells, o);
        if (lane == 0 && my_cells) atomicAdd(A.cells_total, my_cells);
    }
}

size_t k1p_smem_bytes(int qmax, int nqw_max)
{
    return (size_t)K1P_HDR_BYTES + (size_t)2 * (size_t)(nqw_max + K1_QS_EXTRA) * K1P_NT * 4u +
           (size_t)(qmax + 1 + K1_EH_SLACK) * K1P_NT * 8u;
}

template <int SYM>
static cudaError_t k1p_launch_t(const LaunchArgs& a, cudaStream_t st)
{
    if (!a.ntiles) return cudaSuccess;
    if (a.ntiles & 1u) return cudaErrorInvalidValue;               // tiles come in (A, B) pairs
    const size_t smem = k1p_smem_bytes(a.qmax, a.nqw_max);
    auto kern = k1p_extend_kernel<SYM>;
    if (smem > 232448) return cudaErrorInvalidValue;
    static std::atomic<unsigned> smem_set{ 0u };          // per instantiation of this launcher
    cudaError_t err = ensure_max_smem(kern, smem_set);
    if (err != cudaSuccess) return err;
    kern<<<a.ntiles / 2, K1P_NT, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t k1p_launch(const LaunchArgs& a, int sym, cudaStream_t st)
{
    return sym ? k1p_launch_t<1>(a, st) : k1p_launch_t<0>(a, st);
}

}  // namespace bsw
