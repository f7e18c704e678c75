// bsw_device.cuh -- data layout in HBM shared by the host driver and the kernels.
//
// The reference moves a 256 KiB task batch into a BRAM (tbb.v:163-194) that task_parse walks word by
// word (sw_pe_array_task_parse.v:924-948).  Here a batch is re-laid-out by the host scheduler into
// "tiles" so that every device access is a coalesced 128-byte line:
//
//   K1 tile  = 32 extension tasks of similar shape, one per lane of a warp.
//              query  words: q[k*32 + lane], k < nqw     (8 bases per u32, base j in bits 4*(j&7)..+3)
//              target words: t[k*32 + lane], k < ntw
//              so "word k of all 32 tasks" is one 128 B line, the whole query block is one contiguous
//              range that a single TMA bulk copy drops into shared memory in exactly the layout the
//              kernel uses (qs[k][lane]), and the target streams with one coalesced LDG per 8 rows.
//   K2 tile  = 1 long task handled by a whole warp; query and target are stored one base per byte,
//              contiguous, so lane l reads byte j0+l of a row: one coalesced 32-byte sector.
//
// Per-slot scalars are one int4 {qlen, tlen, h0, w} (w already clamped by ksw_extend2's
// max_ins/max_del rule -- the RTL also receives it precomputed: proc_element.v:924-934); results are
// two int4 per slot {score,qle,tle,gtle} {gscore,max_off,cells,status}; the host maps slot -> task.
#pragma once
#include <cstdint>

namespace bsw {

constexpr int TILE_LANES = 32;
constexpr int K1_EH_SLACK = 8;      // row-buffer words past eh[qmax] that a partial chunk may over-read

struct __attribute__((aligned(16))) TileHdr {
    uint32_t qoff16;      // offset of the tile's query block in the sequence arena, in 16-byte units
    uint32_t toff16;      // offset of the tile's target block, in 16-byte units
    uint32_t nqw_ntw;     // K1: words per lane, query (low 16) | target (high 16).  K2: unused
    uint32_t slot0;       // first slot of the tile (K1: 32 slots, K2: 1 slot)
};

struct __attribute__((aligned(16))) SlotParam { int32_t qlen, tlen, h0, w; };   // qlen == 0: padding lane

// Where a slot's packed query / target live in the task-major SOURCE arena (16-byte units).  The host packs every
// sequence once, in input order (a streaming pass); the k0 gather kernel re-lays K1 tiles on the device.
struct SlotSrc { uint32_t qoff16, toff16; };

struct __attribute__((aligned(16))) SlotResult {       // two 16-byte stores per task
    int32_t score, qle, tle, gtle;                      // sw_extend return order (sw_pe_array_sw_extend.v:117-123)
    int32_t gscore, max_off, cells, status;
};

// Scoring parameters (per batch).
struct DevParams {
    int32_t o_del, e_del, o_ins, e_ins, zdrop;
    int32_t match, mismatch;       // FAST scoring: +match / -mismatch (mismatch stored positive)
    uint32_t row_lo[5], row_hi[5]; // GENERIC scoring: row t of the 5x5 matrix as bytes {s(t,0..3)} / {s(t,4),0,0,0}
    int8_t  mat[28];               // the 5x5 matrix itself (K2 GENERIC lookup), padded
    int32_t max_mat;               // max entry of the matrix (ksw_extend2's band clamp, used by the fused seed kernel)
    uint32_t zero;                 // always 0: a zero the compiler cannot fold (keeps it in one register, see bsw_k1_core.cuh)
};

// k0: gather launch = K1 tiles [0, ntiles) of `tiles`; copies src words of every lane into the tiled arena.
struct GatherArgs {
    const TileHdr*   tiles;
    const SlotParam* slots;
    const SlotSrc*   slot_src;
    const uint32_t*  src;          // task-major source arena (H2D copy of the host packer's output)
    uint32_t*        dst;          // tiled arena
    uint32_t         ntiles;
    // raw mode (the caller's bases were copied as they are, one code per byte, from registered host memory): SlotSrc then
    // holds BYTE offsets into raw_q / raw_t, the gather packs the nibbles itself and reports per slot what it saw
    const uint8_t*   raw_q;
    const uint8_t*   raw_t;
    uint32_t*        slot_flags;   // per slot: SLOT_HAS_N | SLOT_BAD_CODE
    const uint32_t*  src2;         // lean path, 2 bit per base: SlotSrc then holds WORD offsets into src2 (16 bases per word)
    TileHdr*         dp_tiles;     // device-planned chunk: the gather takes the tiles' word counts (warp max) into the headers
};

// Device-side scheduler (bsw_plan.cu): the chunk's tasks in input order in, tile / slot arrays out.
struct DpArgs {
    const SlotParam* task_param;   // [count] {qlen, tlen, h0, w} in task order
    const SlotSrc*   task_src;     // [count] where each task's sequences sit in the source arena (or raw byte offsets)
    const uint8_t*   task_cls;     // [count] bit0 = needs matrix-lookup scoring; null: every task has class const_cls
    uint32_t const_cls;
    uint32_t count;
    uint32_t class_count[2];       // tasks per matrix class (host-known: the packer classifies)
    uint32_t class_pos0[2];        // first sorted position of the class (0, class_count[0])
    uint32_t class_slot0[2];       // first slot of the class (a multiple of 32)
    uint32_t class_tile0[2];       // first tile of the class
    uint32_t ntiles;
    uint32_t nmajor;               // non-empty (class, qlen/16) buckets, in sorted order (class ascending, qlen descending)
    uint32_t major_start[192];     // first sorted position of each
    uint8_t  major_of[256];        // (class << 7 | qlen/16) -> index of its bucket
    uint32_t* bins;                // [nmajor * 4096] counters, then running positions
    uint32_t* task_bin;            // [count] bin of every task
    TileHdr*   tiles;              // [ntiles] qoff16 / toff16 / slot0 pre-filled by the host; nqw_ntw written by the K0 gather
    SlotParam* slots;              // [nslots]
    SlotSrc*   slot_src;           // [nslots]
    uint32_t*  out_index;          // [nslots] task index of the slot, 0xffffffff = padding lane
};

// Per-seed scalars of the fused level-2 kernel K3 (one FPGA PE task: sw_pe_array_proc_element.v:1593-1685).
struct __attribute__((aligned(16))) SeedParam {
    int32_t init_score, qbeg, h0;  // regScore, qBeg_ori, h0 (param words 3,4: proc_element.v:871-874,826-828)
    uint32_t id;                   // param word 7 (proc_element.v:807)
    int32_t max_ins[2], max_del[2];// param words 5,6 when they come from the wire (proc_element.v:924-934); -1 = ksw_extend2's formula
};

// One kernel launch = tiles [0, ntiles) of `tiles`.
struct LaunchArgs {
    const TileHdr*   tiles;
    const SlotParam* slots;
    const uint32_t*  arena;        // K1: tiled arena; K2: source arena (16-byte aligned blocks)
    SlotResult*      out;          // indexed by out_index[slot] (the chunk's task order), or by slot when out_index is null
    const uint32_t*  out_index;
    const uint32_t*  slot_flags;   // raw mode: what the gather saw in the slot's bases (null otherwise)
    unsigned long long* cells_total;   // device counter of evaluated DP cells (one atomic per warp), may be null
    DevParams        p;
    uint32_t         ntiles;
    int32_t          qmax;         // max qlen over the launch (sizes the per-lane row buffer)
    int32_t          nqw_max;      // max query words per lane over the launch (K1)
    int32_t          wmax;         // max band over the launch (K2: decides whether the row buffer may be a ring)
    int32_t          k2_narrow;    // K2: rows narrower than 64 columns take the register-only path (option "k2_narrow", default 1)
    // lean output (flat batches planned on the device): the final 24-byte record {score,qle,tle,gtle,gscore,max_off} at
    // out24[out_index[slot]] -- task order, ready for one D2H into the caller's array -- the cell count in cells_out (may
    // be null) and every task with a non-zero status appended to flag_list as (status << 28 | task)
    int32_t*  out24;
    uint32_t* cells_out;
    uint32_t* flag_list;           // [0] = number of entries, entries from [1]
    uint32_t  flag_cap;
    // K3 (fused seed task) only: tiles come in (left, right) pairs, seeds[pair*32 + lane]
    const SeedParam* seeds;
    int32_t          w, pen_clip5, pen_clip3;
};

// K4 (banded global alignment with traceback, ksw_global2)
struct GlobalTask { uint32_t qoff, toff; int32_t qlen, tlen, w; };
struct GlobalArgs {
    const GlobalTask* tasks;
    const uint8_t* qbuf;           // concatenated queries / targets, one base code per byte
    const uint8_t* tbuf;
    int32_t*  eh;                  // row state workspace: tile t starts at eh_off[t] * 32 ints
    uint8_t*  z;                   // direction bytes:     tile t starts at z_off[t] * 32 bytes
    const uint32_t* eh_off;
    const uint64_t* z_off;
    int32_t*  score;
    int32_t*  n_cigar;             // -1: more than max_ops operations
    uint32_t* cigar;               // [ntasks * max_ops]
    uint32_t  ntasks;
    int32_t   max_ops;
    DevParams p;
};

// K5 (wide extension: 32-bit row state in global memory, one warp per task)
struct WideTask { uint64_t row_off; uint32_t qoff, toff; int32_t qlen, tlen, h0, w; };   // row_off: int32 units into `rows`; 2 * (qlen + 2) of them
struct WideArgs {
    const WideTask* tasks;
    const uint8_t* qbuf;           // concatenated queries / targets, one base code per byte
    const uint8_t* tbuf;
    int32_t* rows;
    SlotResult* out;               // [ntasks]
    uint32_t ntasks;
    DevParams p;
};

constexpr int STATUS_OK = 0;
constexpr int STATUS_HAS_N = 2;     // raw mode: the task holds an N and ran on the +a/-b kernel; the host reruns it with matrix lookup
constexpr int STATUS_BAD_CODE = 3;  // raw mode: a base code above 4
constexpr uint32_t SLOT_HAS_N = 1u, SLOT_BAD_CODE = 2u;

}  // namespace bsw
