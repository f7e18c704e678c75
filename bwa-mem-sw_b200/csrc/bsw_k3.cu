// bsw_k3.cu -- K3: fused seed-task kernel (level 2 on the device), one seed per lane, 32 seeds per CTA (sm_100a).
//
// One CTA = one warp = one pair of K1 tiles: the "left" tile holds the 32 seeds' reversed left flanks, the "right" tile
// their right flanks.  Both packed query blocks arrive by TMA bulk copies on one mbarrier; every lane then runs the whole
// processing-element program of its seed (bsw_k3_core.cuh): no host round trip between the left and the right
// extension, band retries in place.  Replaces sw_pe_array_proc_element.v + the two sw_extend calls it makes.
#include <cuda_runtime.h>
#include "bsw_device.cuh"
#include "bsw_k3_core.cuh"
#include "bsw_kernels.h"

namespace bsw {

constexpr int K3_NT = TILE_LANES;
constexpr int K3_HDR_BYTES = 128;

__device__ __forceinline__ uint32_t k3_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int VARIANT, int GENERIC, int SYM>
__global__ void __launch_bounds__(K3_NT) k3_seed_kernel(const __grid_constant__ LaunchArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const TileHdr hl = A.tiles[2 * blockIdx.x], hr = A.tiles[2 * blockIdx.x + 1];
    const uint32_t nql = hl.nqw_ntw & 0x7fffu, nqr = hr.nqw_ntw & 0x7fffu;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    const size_t qwords = (size_t)(A.nqw_max + K1_QS_EXTRA) * K3_NT;
    uint32_t* qsl = reinterpret_cast<uint32_t*>(smem_raw + K3_HDR_BYTES);
    uint32_t* qsr = qsl + qwords;
    uint32_t* eh = qsr + qwords;
    const uint32_t bytes_l = nql * K3_NT * 4u, bytes_r = nqr * K3_NT * 4u;

    if (lane == 0) {
        const uint32_t bar = k3_smem_u32(mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes_l + bytes_r) : "memory");
        if (bytes_l) {
            const void* src = reinterpret_cast<const uint4*>(A.arena) + hl.qoff16;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(k3_smem_u32(qsl)), "l"(src), "r"(bytes_l), "r"(bar) : "memory");
        }
        if (bytes_r) {
            const void* src = reinterpret_cast<const uint4*>(A.arena) + hr.qoff16;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(k3_smem_u32(qsr)), "l"(src), "r"(bytes_r), "r"(bar) : "memory");
        }
    }
    const uint32_t seed_ix = blockIdx.x * K3_NT + lane;
    const SlotParam spl = A.slots[hl.slot0 + lane], spr = A.slots[hr.slot0 + lane];
    const SeedParam sd = A.seeds[seed_ix];
    __syncwarp();
    {
        const uint32_t bar = k3_smem_u32(mbar);
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar) : "memory");
        }
    }
    unsigned long long my_cells = 0;
    if (sd.h0 >= 0) {                                     // h0 < 0 marks a padding lane
        SeedRecord rec;
        uint32_t cells = 0;
        k3_seed<VARIANT, GENERIC, SYM>(A.p, A.w, A.pen_clip5, A.pen_clip3, spl, spr, sd, (int)nql, (int)nqr, eh + lane,
                                       qsl + lane, qsr + lane, A.arena + (size_t)hl.toff16 * 4u + lane,
                                       A.arena + (size_t)hr.toff16 * 4u + lane, rec, cells);
        int4* o = reinterpret_cast<int4*>(A.out + seed_ix);
        o[0] = make_int4((int)rec.id, rec.qb, rec.qe, rec.rb);
        o[1] = make_int4(rec.re, rec.score, rec.truesc, rec.w);
        my_cells = cells;
    }
    if (A.cells_total) {
        for (int o = 16; o; o >>= 1) my_cells += __shfl_xor_sync(0xffffffffu, my_cells, o);
        if (lane == 0 && my_cells) atomicAdd(A.cells_total, my_cells);
    }
}

size_t k3_smem_bytes(int qmax, int nqw_max)
{
    return (size_t)K3_HDR_BYTES + ((size_t)2 * (size_t)(nqw_max + K1_QS_EXTRA) + (size_t)(qmax + 1 + K1_EH_SLACK)) * K3_NT * 4u;
}

template <int VARIANT, int GENERIC, int SYM>
static cudaError_t k3_launch_t(const LaunchArgs& a, cudaStream_t st)
{
    if (!a.ntiles) return cudaSuccess;
    if (a.ntiles & 1u) return cudaErrorInvalidValue;
    const size_t smem = k3_smem_bytes(a.qmax, a.nqw_max);
    auto kern = k3_seed_kernel<VARIANT, GENERIC, SYM>;
    if (smem > 232448) return cudaErrorInvalidValue;
    static std::atomic<unsigned> smem_set{ 0u };          // per instantiation of this launcher
    cudaError_t err = ensure_max_smem(kern, smem_set);
    if (err != cudaSuccess) return err;
    kern<<<a.ntiles / 2, K3_NT, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t k3_launch(const LaunchArgs& a, int variant, int generic, int sym, cudaStream_t st)
{
#define BSW_K3_DISPATCH(V, G, S) if (variant == V && generic == G && sym == S) return k3_launch_t<V, G, S>(a, st);
    BSW_K3_DISPATCH(1, 0, 1) BSW_K3_DISPATCH(1, 0, 0) BSW_K3_DISPATCH(1, 1, 1) BSW_K3_DISPATCH(1, 1, 0)
    BSW_K3_DISPATCH(2, 0, 1) BSW_K3_DISPATCH(2, 0, 0) BSW_K3_DISPATCH(2, 1, 1) BSW_K3_DISPATCH(2, 1, 0)
#undef BSW_K3_DISPATCH
    return cudaErrorInvalidValue;
}

}  // namespace bsw
