// bsw_sched.h -- host-side length-bucketed task scheduler and sequence packer.
//
// Plays the role of the host code that fills the task batch buffer (layout in SURVEY.md App. A.1) and of
// sw_pe_array_task_parse (sw_pe_array_task_parse.v:1600-1650,1652-1762: parse the batch, hand every task to a PE).
// Pure host code, no CUDA: it is unit-tested on the CPU through the emulation harness and used unchanged by
// the product library.  Per chunk of tasks:
//   1. pack_tasks   one streaming pass over the caller's bases: validate the codes, detect N, pack 4 bit/base into the
//                   TASK-MAJOR source arena (each sequence 16-byte aligned) -- sequential reads, sequential writes;
//   2. build_plan   radix-sort the chunk by (class, qlen, tlen, h0), cut it into 32-task tiles and into launches
//                   bucketed by shared-memory occupancy; K1 tiles get a block in the TILED arena, which the device
//                   fills from the source arena (k0 gather kernel, bsw_k0.cu); K2 tiles read the source arena directly.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
#include "bsw_device.cuh"

namespace bsw {

// One extension as the scheduler sees it.  w is the band AFTER the max_ins/max_del clamp.
struct ExtTask {
    const uint8_t* q;
    const uint8_t* t;
    int32_t qlen, tlen, h0, w;
};

struct SchedOptions {
    int variant = 1;            // 1 = RTL / BWA-0.7.8 recurrence, 2 = upstream BWA
    int force_kernel = 0;       // 0 auto, 1 K1 only, 2 K2 only
    int k2_min_qlen = 384;      // auto mode: tasks at least this long go to the intra-task kernel
    int host_threads = 0;       // 0 = hardware concurrency (capped)
    bool fast_matrix = true;    // the 5x5 matrix is (+a / -b, N row/col anything): N-free tasks may use FAST scoring
};

constexpr int K1_QLEN_CAP = 1536;        // shared-memory limit of one K1 tile (227 KB / (32 lanes * 4.5 B per column))
constexpr int K2_QLEN_CAP = 40000;       // shared-memory limit of one K2 task
constexpr int SCORE_CAP   = 32767;       // 16-bit row state: h0 + qlen*max(mat) must not exceed this
// K5 (32-bit rows) takes what K1 / K2 refuse, up to: qlen <= 2^20, tlen <= 2^22, h0 + qlen * (max(mat) + e_ins) and
// max(qlen, tlen) * max(e_del, e_ins) below 2^30 (so that no intermediate of the recurrence leaves int32)
constexpr int WIDE_QLEN_CAP = 1 << 20, WIDE_TLEN_CAP = 1 << 22, WIDE_SCORE_CAP = 0x3fffffff;

struct Launch {
    int kind;          // 1 = K1, 2 = K2, 5 = K3 (pairs: left tile, right tile)
    int generic;       // 1 = matrix lookup scoring
    uint32_t tile0, ntiles;
    int qmax, nqw_max;
    int wmax;          // max band of the launch's tasks
};

struct Plan {
    std::vector<TileHdr>   tiles;        // K1 tiles first (their blocks live in the tiled arena), then K2 tiles
    std::vector<SlotParam> slots;
    std::vector<SlotSrc>   slot_src;     // where each slot's packed query / target sit in the source arena
    std::vector<int64_t>   slot_task;    // task index (into the chunk) of every slot, -1 = padding lane
    std::vector<Launch>    launches;
    uint32_t n_k1_tiles = 0;
    size_t tiled_words = 0;              // size of the tiled arena (device only)
    uint64_t est_cells = 0;
    // scratch reused across calls (radix sort of the chunk)
    std::vector<uint32_t> key, order, tmp, hist;
    // K3 plans: seed index of every (tile pair, lane), -1 = padding lane
    std::vector<int64_t> lane_seed;
};

// Upper bound of the source arena for these tasks, in u32 words (16-byte aligned sequences + slack).
size_t source_arena_bound(const ExtTask* tasks, size_t n);

// Fused validate + classify + pack (single pass, single thread).  cls: bit0 = needs matrix-lookup scoring (contains N,
// or the matrix is not +a/-b), bit1 = long task (K2).  src[i] = offsets of task i's packed sequences in `arena`.
// Returns 0 or a negative BSW_E* code; on error *bad_task is the first offending task and msg explains.
int pack_tasks(const ExtTask* tasks, size_t n, int max_mat, const SchedOptions& opt, uint8_t* cls, SlotSrc* src,
               uint32_t* arena, size_t* words_used, size_t* bad_task, std::string* msg);
// the two builds of bsw_pack.cpp behind pack_tasks (run-time dispatch on the CPU's ISA)
int pack_tasks_sse2(const ExtTask* tasks, size_t n, int max_mat, const SchedOptions& opt, uint8_t* cls, SlotSrc* src,
                    uint32_t* arena, size_t* words_used, size_t* bad_task, std::string* msg);
int pack_tasks_avx512(const ExtTask* tasks, size_t n, int max_mat, const SchedOptions& opt, uint8_t* cls, SlotSrc* src,
                      uint32_t* arena, size_t* words_used, size_t* bad_task, std::string* msg);

// 2 bit per base packer of the lean flat path (bsw_pack2.cpp, two builds; pack2_flat dispatches on the CPU).
int64_t pack2_flat_generic(const uint8_t* qbuf, const int64_t* qoff, const uint8_t* tbuf, const int64_t* toff, size_t count,
                           uint32_t* arena, SlotSrc* src, std::vector<uint32_t>* n_list);
int64_t pack2_flat_avx512(const uint8_t* qbuf, const int64_t* qoff, const uint8_t* tbuf, const int64_t* toff, size_t count,
                          uint32_t* arena, SlotSrc* src, std::vector<uint32_t>* n_list);
int64_t pack2_flat(const uint8_t* qbuf, const int64_t* qoff, const uint8_t* tbuf, const int64_t* toff, size_t count,
                   uint32_t* arena, SlotSrc* src, std::vector<uint32_t>* n_list);

// Level-2 plan: tasks[2*s] / tasks[2*s+1] are the left / right flank of seed s (qlen == 0: absent, cls 0x80).  The band
// of a seed task is decided on the device, so ExtTask.w is free: a present right flank carries its score-budget hint
// there (about h0 + left qlen), which only feeds the sort key.  Seeds are
// sorted by their longer flank and cut into pairs of K1 tiles (left flanks, right flanks), 32 seeds per pair.
void build_seed_plan(const ExtTask* tasks, const uint8_t* cls, const SlotSrc* src, size_t nseeds, const SchedOptions& opt, Plan* plan);

// Sort, tile, bucket.  Single-threaded: the driver runs one plan per chunk per host thread.
void build_plan(const ExtTask* tasks, const uint8_t* cls, const SlotSrc* src, size_t n, const SchedOptions& opt, Plan* plan);

// Device-side scheduling (bsw_plan.cu sorts and builds the slots): the host only needs the launch geometry, which
// follows from per-bucket counts -- the sort key's major field is qlen/16, so class c's tiles are runs of buckets in
// descending order.  Fills plan->tiles (offsets and slot0; word counts are upper bounds the device replaces),
// plan->launches, n_k1_tiles, tiled_words.  Returns false when the chunk holds a task the device planner does not take
// (long tasks: the K2 classes), in which case build_plan must be used.
struct DpGeometry {
    uint32_t class_count[2], class_pos0[2], class_slot0[2], class_tile0[2]; uint32_t ntiles; size_t nslots;
    uint32_t nmajor, major_start[192]; uint8_t major_of[256];      // non-empty (class, qlen/16) buckets in sorted order
};
bool build_dp_plan(const ExtTask* tasks, const uint8_t* cls, size_t n, const SchedOptions& opt, Plan* plan, DpGeometry* g);
// The two halves of build_dp_plan for callers that walk their tasks themselves (the lean flat path of bsw_host.cpp).
struct DpBuckets {
    struct B { uint32_t cnt; int maxq, maxt, maxw; } bk[2][128];         // [matrix class][qlen / 16]
    void add(int cls, int qlen, int tlen, int w)
    {
        B& b = bk[cls & 1][qlen >> 4 < 127 ? qlen >> 4 : 127];
        ++b.cnt;
        if (qlen > b.maxq) b.maxq = qlen;
        if (tlen > b.maxt) b.maxt = tlen;
        if (w > b.maxw) b.maxw = w;
    }
};
bool dp_geometry(const DpBuckets& bks, size_t n, const SchedOptions& opt, Plan* plan, DpGeometry* g);

// ksw_extend2's band clamp (public BWA algorithm; the RTL takes max_ins/max_del precomputed from the host:
// sw_pe_array_proc_element.v:924-934 and applies them at sw_pe_array_sw_extend.v:1763-1765,1881,1890).
int clamp_band(const int8_t mat[25], int qlen, int w, int end_bonus, int o_ins, int e_ins, int o_del, int e_del);

// The same clamp with the two divisions tabulated per query length (built once per call; short queries are the bulk).
struct BandClamp {
    static constexpr int TABLE = 1024;
    int8_t mat[25]; int end_bonus, o_ins, e_ins, o_del, e_del;
    int bound[TABLE + 1];
    BandClamp(const int8_t m[25], int end_bonus_, int o_ins_, int e_ins_, int o_del_, int e_del_)
        : end_bonus(end_bonus_), o_ins(o_ins_), e_ins(e_ins_), o_del(o_del_), e_del(e_del_)
    {
        for (int k = 0; k < 25; ++k) mat[k] = m[k];
        for (int q = 0; q <= TABLE; ++q) bound[q] = clamp_band(mat, q, 0x7fffffff, end_bonus, o_ins, e_ins, o_del, e_del);
    }
    int operator()(int qlen, int w) const
    {
        if (qlen >= 0 && qlen <= TABLE) return w < bound[qlen] ? w : bound[qlen];
        return clamp_band(mat, qlen, w, end_bonus, o_ins, e_ins, o_del, e_del);
    }
};

// Simple fork-join helper.
void parallel_for(size_t n, size_t grain, int nthreads, const void* ctx,
                  void (*fn)(const void* ctx, size_t lo, size_t hi));
int default_host_threads();

template <class F>
inline void pfor(size_t n, size_t grain, int nthreads, F&& f)
{
    auto tramp = [](const void* c, size_t lo, size_t hi) { (*reinterpret_cast<const F*>(c))(lo, hi); };
    parallel_for(n, grain, nthreads, &f, tramp);
}

}  // namespace bsw
