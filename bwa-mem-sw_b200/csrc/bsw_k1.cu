// bsw_k1.cu -- K1: inter-task extension kernel, one thread per extension task (sm_100a).
//
// Replaces the 80 free-running PEs of the reference (sw_pe_array.v:1133-1494, one task per PE,
// sw_pe_array_sw_extend.v FSM :1639-1705) by thousands of resident lanes: one CTA = one warp = one
// tile of 32 tasks of similar shape (the host scheduler sorts by length, bsw_sched.cpp), the grid is
// the tile list in longest-first order, and the hardware CTA scheduler plays the role of task_parse's
// "next PE with room" dispatch (sw_pe_array_task_parse.v:1600-1650).
//
// Data movement: the tile's packed query block is contiguous in HBM in exactly the shared-memory
// layout (word k of lane l at [k*32+l]); one elected lane drops it into shared memory with a single TMA
// bulk copy (cp.async.bulk -> UBLKCP) tracked by an mbarrier while the other lanes fetch their slot
// scalars.  The target is streamed from HBM with one coalesced 128-byte load per 8 rows, prefetched one
// group ahead.  The DP itself is in bsw_k1_core.cuh.
#include <cuda_runtime.h>
#include "bsw_device.cuh"
#include "bsw_k1_core.cuh"
#include "bsw_kernels.h"

namespace bsw {

constexpr int K1_NT = TILE_LANES;   // one warp per CTA: no block-level synchronisation anywhere in K1
constexpr int K1_HDR_BYTES = 128;   // mbarrier + padding in front of the query block

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int VARIANT, int GENERIC, int SYM>
__global__ void __launch_bounds__(K1_NT) k1_extend_kernel(const __grid_constant__ LaunchArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    const TileHdr hd = A.tiles[blockIdx.x];
    const uint32_t nqw = hd.nqw_ntw & 0x7fffu;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* qs = reinterpret_cast<uint32_t*>(smem_raw + K1_HDR_BYTES);
    uint32_t* eh = qs + (size_t)(A.nqw_max + K1_QS_EXTRA) * K1_NT;
    const uint32_t qbytes = nqw * K1_NT * 4u;

    if (lane == 0) {
        const uint32_t bar = smem_u32(mbar), dst = smem_u32(qs);
        const void* src = reinterpret_cast<const uint4*>(A.arena) + hd.qoff16;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(qbytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src), "r"(qbytes), "r"(bar) : "memory");
    }
    const uint32_t slot = hd.slot0 + lane;
    const SlotParam sp = A.slots[slot];
    __syncwarp();
    {
        const uint32_t bar = smem_u32(mbar);
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar) : "memory");
        }
    }

    unsigned long long my_cells = 0;
    if (sp.qlen > 0) {
        const uint32_t* tg = A.arena + (size_t)hd.toff16 * 4u + lane;
        SlotResult r;
        k1_task<VARIANT, GENERIC, SYM>(A.p, sp.qlen, sp.tlen, sp.h0, sp.w, (int)nqw, eh + lane, qs + lane, tg, r);
        if (A.slot_flags) {                              // raw mode: what the gather saw in this slot's bases
            const uint32_t f = A.slot_flags[slot];
            if (f & SLOT_BAD_CODE) r.status = STATUS_BAD_CODE;
            else if (!GENERIC && (f & SLOT_HAS_N)) r.status = STATUS_HAS_N;      // the +a/-b cell cannot score an N
        }
        if (A.out24) {
            const uint32_t task = A.out_index[slot];
            int2* o = reinterpret_cast<int2*>(A.out24 + (size_t)task * 6);
            o[0] = make_int2(r.score, r.qle);
            o[1] = make_int2(r.tle, r.gtle);
            o[2] = make_int2(r.gscore, r.max_off);
            if (A.cells_out) A.cells_out[task] = (uint32_t)r.cells;
            if (r.status != STATUS_OK) {
                const uint32_t k = atomicAdd(A.flag_list, 1u);
                if (k < A.flag_cap) A.flag_list[1 + k] = ((uint32_t)r.status << 28) | task;
            }
        } else {
            int4* o = reinterpret_cast<int4*>(A.out + (A.out_index ? A.out_index[slot] : slot));
            o[0] = make_int4(r.score, r.qle, r.tle, r.gtle);
            o[1] = make_int4(r.gscore, r.max_off, r.cells, r.status);
        }
        my_cells = (uint32_t)r.cells;
    }
    if (A.cells_total) {                                 // one atomic per warp for the device-side cell counter
        for (int o = 16; o; o >>= 1) my_cells += __shfl_xor_sync(0xffffffffu, my_cells, o);
        if (lane == 0 && my_cells) atomicAdd(A.cells_total, my_cells);
    }
}

size_t k1_smem_bytes(int qmax, int nqw_max)
{
    return (size_t)K1_HDR_BYTES + ((size_t)(nqw_max + K1_QS_EXTRA) + (size_t)(qmax + 1 + K1_EH_SLACK)) * K1_NT * 4u;
}

template <int VARIANT, int GENERIC, int SYM>
static cudaError_t k1_launch_t(const LaunchArgs& a, cudaStream_t st)
{
    if (!a.ntiles) return cudaSuccess;
    const size_t smem = k1_smem_bytes(a.qmax, a.nqw_max);
    auto kern = k1_extend_kernel<VARIANT, GENERIC, SYM>;
    // always the same (maximal) value: launches are issued concurrently from several host threads, and a per-launch
    // value would race with another thread's launch of the same kernel
    if (smem > 232448) return cudaErrorInvalidValue;
    static std::atomic<unsigned> smem_set{ 0u };          // per instantiation of this launcher
    cudaError_t err = ensure_max_smem(kern, smem_set);
    if (err != cudaSuccess) return err;
    kern<<<a.ntiles, K1_NT, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t k1_launch(const LaunchArgs& a, int variant, int generic, int sym, cudaStream_t st)
{
#define BSW_K1_DISPATCH(V, G, S) if (variant == V && generic == G && sym == S) return k1_launch_t<V, G, S>(a, st);
    BSW_K1_DISPATCH(1, 0, 1) BSW_K1_DISPATCH(1, 0, 0) BSW_K1_DISPATCH(1, 1, 1) BSW_K1_DISPATCH(1, 1, 0)
    BSW_K1_DISPATCH(2, 0, 1) BSW_K1_DISPATCH(2, 0, 0) BSW_K1_DISPATCH(2, 1, 1) BSW_K1_DISPATCH(2, 1, 0)
#undef BSW_K1_DISPATCH
    return cudaErrorInvalidValue;
}

}  // namespace bsw
