// bsw_device.cuh -- device-side data layout shared by the host driver and the kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bsw {

// Per-chunk task table in HBM (structure of arrays, indexed by task id inside the chunk).
//   qseq/tseq : bases packed 4 bit each, base j of a sequence in word j>>3, bits 4*(j&7)..+3
//               (little-nibble first; the FPGA wire order -- first base in bits 31:28,
//               proc_element.v:1638,1677 -- is converted by the level-3 adapter).  Every
//               sequence starts on a 16-byte boundary and is zero-padded to a multiple of 16 bytes,
//               and both arrays end with 32 bytes of slack, so kernels may over-read one uint4.
//   qoffw/toffw : word offset of the task's first query/target word.
//   w          : band width AFTER ksw_extend2's max_ins/max_del clamp (done on the host in double,
//                exactly as BWA does; the RTL also receives it precomputed: proc_element.v:924-934).
struct DevTasks {
    const uint32_t* qseq;
    const uint32_t* tseq;
    const uint32_t* qoffw;
    const uint32_t* toffw;
    const int32_t*  qlen;
    const int32_t*  tlen;
    const int32_t*  h0;
    const int32_t*  w;
    int4*           out;      // 2 x int4 per task: {score,qle,tle,gtle} {gscore,max_off,cells,status}
};

// Scoring parameters (per batch).
struct DevParams {
    int32_t o_del, e_del, o_ins, e_ins, zdrop;
    int32_t match, mismatch;      // FAST scoring: +match / -mismatch (mismatch stored positive), valid iff fast_ok
    uint32_t row_lo[5], row_hi[5];// GENERIC scoring: row t of the 5x5 matrix as bytes {s(t,0..3)} / {s(t,4),0,0,0}
};

// One kernel launch = a slice [slot0, slot1) of `order` (task ids sorted by the scheduler).
struct LaunchArgs {
    DevTasks  t;
    DevParams p;
    const uint32_t* order;
    uint32_t slot0, slot1;
    int32_t  qmax;            // max qlen in the slice (sizes the per-thread row buffer)
    unsigned long long* cells_total;   // device counter (atomicAdd once per warp)
};

constexpr int STATUS_OK = 0;

}  // namespace bsw
