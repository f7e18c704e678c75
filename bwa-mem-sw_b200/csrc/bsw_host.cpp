// bsw_host.cpp -- libbsw.so: context, pinned staging, streams, level-1/2 batch calls (C ABI in include/bsw.h).
//
// Replaces the reference's transport + control plane for this path: the AAL/CCI session and CSR writes
// (batch_manager.v:208-213,313-351), the task/result batch buffers (tbb.v, rbb.v) and the DSM busy-bit polling
// (batch_manager.v:851-854) become pinned host buffers, cudaMemcpyAsync on per-device CUDA streams and events.
// There is no CPU fallback: without a CUDA device bsw_init fails, and every batch call runs the sm_100a kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <deque>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bsw.h"
#include "bsw_device.cuh"
#include "bsw_kernels.h"
#include "bsw_sched.h"
#include "bsw_internal.h"

using namespace bsw;

namespace {

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

constexpr uint32_t LEAN_FLAGS = 1023;      // flagged tasks a lean chunk reports in its first D2H (more: the list is fetched again)

// One staging slot = one in-flight chunk on one stream.
struct Slot {
    cudaStream_t stream = nullptr;
    cudaStream_t side[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };   // extra streams: the bucket launches of one big plan overlap their tails
    cudaEvent_t ev_fork = nullptr, ev_join[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_done = nullptr, ev_in = nullptr, ev_out = nullptr;   // ev_in/ev_out: BSW_TRACE only
    // One pinned input block per chunk, mirrored on the device and moved with a single cudaMemcpyAsync:
    //   [ source arena (task-major packed sequences) | TileHdr[] | SlotParam[] | SlotSrc[] | u32 task-of-slot[] ]   (16-byte aligned parts)
    unsigned char* h_in = nullptr; size_t h_in_cap = 0;
    unsigned char* d_in = nullptr; size_t d_in_cap = 0;
    size_t in_bytes = 0, off_tiles = 0, off_slots = 0, off_ssrc = 0, off_oidx = 0;
    SlotResult* h_out = nullptr;  size_t h_out_cap = 0;
    uint32_t* d_arena = nullptr;  size_t d_arena_cap = 0;     // tiled arena, written by the k0 gather kernel
    SlotResult* d_out = nullptr;  size_t d_out_cap = 0;
    size_t src_words = 0;
    // raw mode: the chunk's bases as the caller holds them (one code per byte) and the gather's per-slot findings
    unsigned char* d_rawq = nullptr; size_t d_rawq_cap = 0;
    unsigned char* d_rawt = nullptr; size_t d_rawt_cap = 0;
    uint32_t* d_flags = nullptr;     size_t d_flags_cap = 0;
    bool raw_mode = false;
    // device-side scheduling (bsw_plan.cu): the input block then carries the tasks in input order (off_slots / off_ssrc
    // / off_oidx = task params / task sources / task classes) and the slot arrays live in d_plan, written by the device
    bool dp_mode = false;
    unsigned char* d_plan = nullptr; size_t d_plan_cap = 0;
    size_t dp_off_ssrc = 0, dp_off_oidx = 0;
    // lean mode (flat batch, registered bases, device planner): 24-byte records in task order, straight into the
    // caller's array when that is page-locked too; flags = [count | entries...] of tasks with a non-zero status
    bool lean = false, lean_direct = false, lean_cells = false;
    size_t off_arena2 = 0;         // lean path with host-packed 2-bit bases: offset of the packed arena in the input block (0: raw bases)
    int32_t* d_out24 = nullptr; size_t d_out24_cap = 0;      // 6 x int32 per task
    int32_t* h_out24 = nullptr; size_t h_out24_cap = 0;
    uint32_t* d_cellsv = nullptr; size_t d_cellsv_cap = 0;
    uint32_t* h_cellsv = nullptr; size_t h_cellsv_cap = 0;
    uint32_t* d_flaglist = nullptr; uint32_t* h_flaglist = nullptr;      // LEAN_FLAGS + 1 words each
    const uint32_t*  d_src() const { return reinterpret_cast<const uint32_t*>(d_in); }
    const TileHdr*   d_tiles() const { return reinterpret_cast<const TileHdr*>(d_in + off_tiles); }
    const SlotParam* d_slots() const { return dp_mode ? reinterpret_cast<const SlotParam*>(d_plan) : reinterpret_cast<const SlotParam*>(d_in + off_slots); }
    const SlotSrc*   d_ssrc() const { return dp_mode ? reinterpret_cast<const SlotSrc*>(d_plan + dp_off_ssrc) : reinterpret_cast<const SlotSrc*>(d_in + off_ssrc); }
    const uint32_t*  d_oidx() const { return dp_mode ? reinterpret_cast<const uint32_t*>(d_plan + dp_off_oidx) : reinterpret_cast<const uint32_t*>(d_in + off_oidx); }
    unsigned long long* d_cells = nullptr;
    unsigned long long* h_cells = nullptr;
    // in-flight bookkeeping
    bool busy = false, timed = false;
    double trace_ms[3] = { 0, 0, 0 };     // pack, plan, enqueue time of the last submit (BSW_TRACE)
    Plan plan;
    size_t first = 0, count = 0;     // chunk = tasks [first, first+count) of the batch
    size_t nlaunch = 0;
    // per-chunk host scratch (reused)
    std::vector<ExtTask> tasks;
    std::vector<uint8_t> cls;
    std::vector<SlotSrc> src;
    std::vector<size_t> seed_idx;    // level 2: the caller's seed index of every seed of the chunk in flight
};

// One host worker thread = one pipeline: it owns two staging slots (streams) on one device and alternates between
// them, so the chunk it packs overlaps the chunk the GPU is computing.
struct Worker {
    int dev = 0;
    bool ready = false;
    std::vector<Slot> slots;         // stream slots the worker cycles through (bsw_ctx::slots_per_worker)
};

struct Device {
    int id = 0;
    Slot aux;                        // stream for measurement kernels
};

struct FlatSrc {
    const bsw_params* p; const BandClamp* clamp; const uint8_t* qbuf; const int64_t* qoff; const uint8_t* tbuf; const int64_t* toff;
    const int32_t* h0; const int32_t* w;
};

// Where a batch's tasks come from (flat arrays, bsw_task records, or an internal ExtTask vector).
struct TaskSource {
    const void* self;
    void (*fill)(const void* self, size_t first, size_t count, ExtTask* out);
    // raw mode: the bases of consecutive tasks are consecutive in two buffers the caller registered with bsw_host_register
    // (so the DMA engine can read them in place); a chunk's bases are then the byte ranges [task first .q, last .q + qlen)
    bool raw = false;
    const FlatSrc* flat = nullptr;      // set for flat batches: the lean path reads the caller's arrays directly
    bool out_registered = false;               // the caller's result array is page-locked: results are copied straight into it
};

}  // namespace

struct bsw_ctx {
    std::vector<Device> devs;
    // Every batch call in flight owns a CallState (its worker pipelines: streams, pinned and device staging).  The FPGA
    // keeps four batches in flight -- array k computes while k+1 is fetched (batch_manager.v:397-562, tbb.v:110-116) --
    // so concurrent calls on one context must overlap, not queue: a call takes a free CallState (or makes one), runs
    // without any context-wide lock, and puts it back.
    struct CallState { std::vector<std::unique_ptr<Worker>> workers; };
    std::vector<std::unique_ptr<CallState>> call_free, call_all;
    std::mutex pool_mu;
    std::atomic<int> calls_live{ 0 };      // batch calls in flight: they share the host threads instead of each taking all
    int streams_per_device = 2;
    SchedOptions opt;
    size_t chunk_tasks = 16384;
    int slots_per_worker = 3;      // chunks one worker keeps in flight (3 measured 5-20 % faster than 2 once the host work is small)
    int raw_inputs = 2;            // flat batches whose base buffers are registered (bsw_host_register) skip the host packer:
                                   // 0 never, 1 always, 2 auto (when there are at most 10 host threads per GPU)
    std::vector<std::pair<const unsigned char*, size_t>> host_regs;   // registered host ranges      // chunks one worker keeps in flight
    int k2_warps = 1;              // warps per K2 task (1: most tasks per SM; 4: widest rows in parallel)
    int wide = 1;                  // tasks outside the 16-bit envelope of K1 / K2: 0 = refuse the batch (BSW_ERANGE), 1 = run them on K5
                                   // (32-bit rows), 2 = run EVERY task on K5 (tests)
    bool fpga_strict = false;      // bsw_fpga_batch refuses (BSW_ERANGE) batches with a task outside the FPGA's 8-bit envelope
    bool k2_narrow = true;         // K2 rows below 64 columns run in registers (bsw_k2.cu::k2_narrow_row)
    bool device_plan = true;       // the chunk's sort + tile building run on the device (bsw_plan.cu); false: host build_plan
    bool fused_l2 = true;          // level 2 runs as one fused kernel (K3); false: host-orchestrated level-1 passes
    cudaEvent_t trace_ref = nullptr;   // BSW_TRACE: recorded at the start of a batch call, origin of the per-chunk GPU timeline
    double trace_ref_host_ms = 0;      // host time (relative to the call start) at which trace_ref was recorded
    bool kernel_timing = false;    // record CUDA events around each chunk's kernels (bsw_stats.kernel_ms); two more driver calls per chunk
    std::mutex mu;                 // serialises batch calls on this context
    std::mutex err_mu;
    std::string last_error;
    bsw_stats stats{};
    std::mutex async_mu;
    struct Async { uint32_t seq = 0; bool used = false, done = false; int status = 0; std::string error; };
    std::vector<Async> async;
    uint32_t async_seq = 0;
    std::deque<std::pair<size_t, std::function<int()>>> async_queue;
    std::condition_variable async_cv, async_done_cv;
    std::vector<std::thread> async_threads;
    bool async_stop = false;
    void async_main();
};

struct bsw_resident {
    int dev = 0;
    Slot slot;
    DevParams dp{};
    int sym = 0, variant = 1;
    size_t n = 0;
    double last_ms = 0;
};

namespace {

// The last error text is kept per calling thread (several threads may share a context); worker threads of a call
// also write the context-wide copy, which the calling thread picks up when its call returns.
thread_local std::string tl_last_error;

void set_error(bsw_ctx* ctx, const std::string& s)
{
    tl_last_error = s;
    if (!ctx) return;
    std::lock_guard<std::mutex> g(ctx->err_mu);
    ctx->last_error = s;
}

int cuda_fail(bsw_ctx* ctx, cudaError_t e, const char* what)
{
    set_error(ctx, std::string(what) + ": " + cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? BSW_ENOMEM : BSW_ECUDA;
}

#define CUDA_TRY(ctx, call)                                              \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) return cuda_fail((ctx), e__, #call);     \
    } while (0)

template <class T>
int grow_pinned(bsw_ctx* ctx, T** p, size_t* cap, size_t need)
{
    if (need <= *cap) return 0;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *cap = 0;
    // generous headroom: chunks differ in size from call to call, and a pinned reallocation costs milliseconds
    const size_t want = need + need / 2 + 65536 / sizeof(T) + 64;
    CUDA_TRY(ctx, cudaHostAlloc((void**)p, want * sizeof(T), cudaHostAllocDefault));
    *cap = want;
    return 0;
}
template <class T>
int grow_device(bsw_ctx* ctx, T** p, size_t* cap, size_t need)
{
    if (need <= *cap) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr; *cap = 0;
    const size_t want = need + need / 2 + 65536 / sizeof(T) + 64;
    CUDA_TRY(ctx, cudaMalloc((void**)p, want * sizeof(T)));
    *cap = want;
    return 0;
}

int slot_init(bsw_ctx* ctx, Slot& s)
{
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CUDA_TRY(ctx, cudaEventCreate(&s.ev_k0));
    CUDA_TRY(ctx, cudaEventCreate(&s.ev_in));
    CUDA_TRY(ctx, cudaEventCreate(&s.ev_out));
    CUDA_TRY(ctx, cudaEventCreate(&s.ev_k1));
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming | cudaEventBlockingSync));
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&s.ev_fork, cudaEventDisableTiming));
    for (int k = 0; k < 5; ++k) {
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&s.side[k], cudaStreamNonBlocking));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&s.ev_join[k], cudaEventDisableTiming));
    }
    CUDA_TRY(ctx, cudaMalloc((void**)&s.d_cells, sizeof(unsigned long long)));
    CUDA_TRY(ctx, cudaHostAlloc((void**)&s.h_cells, sizeof(unsigned long long), cudaHostAllocDefault));
    return 0;
}

void slot_free(Slot& s)
{
    if (s.h_in) cudaFreeHost(s.h_in);
    if (s.d_in) cudaFree(s.d_in);
    if (s.h_out) cudaFreeHost(s.h_out);
    if (s.h_cells) cudaFreeHost(s.h_cells);
    if (s.d_arena) cudaFree(s.d_arena);
    if (s.d_out) cudaFree(s.d_out);
    if (s.d_cells) cudaFree(s.d_cells);
    if (s.d_rawq) cudaFree(s.d_rawq);
    if (s.d_rawt) cudaFree(s.d_rawt);
    if (s.d_flags) cudaFree(s.d_flags);
    if (s.d_plan) cudaFree(s.d_plan);
    if (s.d_out24) cudaFree(s.d_out24);
    if (s.h_out24) cudaFreeHost(s.h_out24);
    if (s.d_cellsv) cudaFree(s.d_cellsv);
    if (s.h_cellsv) cudaFreeHost(s.h_cellsv);
    if (s.d_flaglist) cudaFree(s.d_flaglist);
    if (s.h_flaglist) cudaFreeHost(s.h_flaglist);
    if (s.ev_k0) cudaEventDestroy(s.ev_k0);
    if (s.ev_in) cudaEventDestroy(s.ev_in);
    if (s.ev_out) cudaEventDestroy(s.ev_out);
    if (s.ev_k1) cudaEventDestroy(s.ev_k1);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    if (s.ev_fork) cudaEventDestroy(s.ev_fork);
    for (int k = 0; k < 5; ++k) { if (s.ev_join[k]) cudaEventDestroy(s.ev_join[k]); if (s.side[k]) cudaStreamDestroy(s.side[k]); }
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Slot();
}

int make_dev_params(bsw_ctx* ctx, const bsw_params* p, DevParams* dp, int* sym, bool* fast_ok, int* max_mat)
{
    if (!p) { set_error(ctx, "params is null"); return BSW_EINVAL; }
    if (p->e_del < 1 || p->e_ins < 1 || p->o_del < 0 || p->o_ins < 0 ||
        p->e_del > 1000 || p->e_ins > 1000 || p->o_del > 10000 || p->o_ins > 10000) {
        set_error(ctx, "gap penalties out of range (need 1 <= e <= 1000, 0 <= o <= 10000)");
        return BSW_EINVAL;
    }
    memset(dp, 0, sizeof(*dp));
    dp->o_del = p->o_del; dp->e_del = p->e_del; dp->o_ins = p->o_ins; dp->e_ins = p->e_ins; dp->zdrop = p->zdrop;
    int mx = 0;
    for (int k = 0; k < 25; ++k) { dp->mat[k] = p->mat[k]; mx = mx > p->mat[k] ? mx : p->mat[k]; }
    *max_mat = mx;
    dp->max_mat = mx;
    // FAST scoring applies to N-free tasks when the 4x4 core is +a on the diagonal and -b elsewhere
    bool fast = true;
    const int a = p->mat[0], b = -p->mat[1];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (p->mat[5 * i + j] != (i == j ? a : -b)) fast = false;
    *fast_ok = fast;
    dp->match = a; dp->mismatch = b;
    for (int t = 0; t < 5; ++t) {
        uint32_t lo = 0;
        for (int q = 0; q < 4; ++q) lo |= (uint32_t)(uint8_t)p->mat[5 * t + q] << (8 * q);
        dp->row_lo[t] = lo;
        dp->row_hi[t] = (uint32_t)(uint8_t)p->mat[5 * t + 4];
    }
    *sym = (p->o_del == p->o_ins && p->e_del == p->e_ins) ? 1 : 0;
    return 0;
}

// Host statistics accumulated by one worker and merged once per call.
struct LocalStats { double pack_ms = 0, validate_ms = 0, kernel_ms = 0; uint64_t h2d = 0, d2h = 0, launches = 0, tasks = 0, cells = 0, packed_chunks = 0, raw_chunks = 0; };

int enqueue_launches(bsw_ctx* ctx, Slot& s, const DevParams& dp, int sym, int variant, bool count_cells, size_t* nlaunch, const uint32_t* out_index)
{
    const Plan& P = s.plan;
    size_t nl = 0;
    // A plan has one launch per class and occupancy bucket; issued on one stream each would wait for the previous
    // launch's last CTA (0.2-0.4 ms of tile latency each, measured 1.0 ms -> 0.38 ms for a 32 k task chunk).  Spread
    // them over the slot's four streams (fork after the gather, join before the D2H).
    static const int nside = getenv("BSW_SIDE_STREAMS") ? std::max(0, std::min(5, atoi(getenv("BSW_SIDE_STREAMS")))) : 3;
    // only as many side streams as there are extra launches (every fork/join is four driver calls)
    const int nuse = (int)std::min<size_t>((size_t)nside, P.launches.size() > 0 ? P.launches.size() - 1 : 0);
    const bool spread = nuse > 0;
    if (spread) {
        CUDA_TRY(ctx, cudaEventRecord(s.ev_fork, s.stream));
        for (int k = 0; k < nuse; ++k) CUDA_TRY(ctx, cudaStreamWaitEvent(s.side[k], s.ev_fork, 0));
    }
    for (const Launch& L : P.launches) {
        LaunchArgs a{};
        a.tiles = s.d_tiles() + L.tile0; a.slots = s.d_slots(); a.arena = (L.kind == 2) ? s.d_src() : s.d_arena; a.out = s.d_out; a.out_index = out_index; a.slot_flags = s.raw_mode ? s.d_flags : nullptr;
        a.cells_total = count_cells ? s.d_cells : nullptr; a.p = dp; a.ntiles = L.ntiles; a.qmax = L.qmax; a.nqw_max = L.nqw_max; a.wmax = L.wmax;
        a.k2_narrow = ctx->k2_narrow ? 1 : 0;
        if (s.lean) { a.out24 = s.d_out24; a.cells_out = s.lean_cells ? s.d_cellsv : nullptr; a.flag_list = s.d_flaglist; a.flag_cap = LEAN_FLAGS; }
        const size_t lane_ix = nl % (size_t)(nuse + 1);
        cudaStream_t st = (spread && lane_ix) ? s.side[lane_ix - 1] : s.stream;
        cudaError_t e = (L.kind == 1) ? k1_launch(a, variant, L.generic, sym, st)
                                      : k2_launch(a, L.generic, ctx->k2_warps, variant, st);
        if (e != cudaSuccess) return cuda_fail(ctx, e, L.kind == 1 ? "K1 launch" : "K2 launch");
        ++nl;
    }
    if (spread) {
        for (int k = 0; k < nuse; ++k) {
            CUDA_TRY(ctx, cudaEventRecord(s.ev_join[k], s.side[k]));
            CUDA_TRY(ctx, cudaStreamWaitEvent(s.stream, s.ev_join[k], 0));
        }
    }
    *nlaunch = nl;
    return 0;
}

int enqueue_gather(bsw_ctx* ctx, Slot& s)
{
    const Plan& P = s.plan;
    if (!P.n_k1_tiles) return 0;
    GatherArgs g{};
    g.tiles = s.d_tiles(); g.slots = s.d_slots(); g.slot_src = s.d_ssrc(); g.src = s.d_src(); g.dst = s.d_arena; g.ntiles = P.n_k1_tiles;
    if (s.raw_mode) { g.raw_q = s.d_rawq; g.raw_t = s.d_rawt; g.slot_flags = s.d_flags; }
    if (s.lean && s.off_arena2) g.src2 = reinterpret_cast<const uint32_t*>(s.d_in + s.off_arena2);
    if (s.dp_mode) g.dp_tiles = reinterpret_cast<TileHdr*>(s.d_in + s.off_tiles);
    cudaError_t e = k0_launch(g, s.stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "K0 gather launch");
    return 0;
}

// Pack + schedule the chunk held in s.tasks (tasks [first, first+count) of the batch) into the slot's pinned staging,
// then enqueue H2D, the k0 gather, the extension kernels and D2H on the slot's stream.
// Raw mode takes a chunk as it is when every task is a plain inter-task (K1) extension with valid scalars; anything else
// (long tasks, bad lengths, ...) goes through the staged path, which also words the error messages.
bool raw_chunk_eligible(const ExtTask* t, size_t count, int max_mat, const SchedOptions& opt)
{
    if (opt.force_kernel == 2 || count == 0) return false;
    const uint8_t* q0 = t[0].q; const uint8_t* t0 = t[0].t;
    if (!q0 || !t0) return false;
    for (size_t i = 0; i < count; ++i) {
        const ExtTask& x = t[i];
        if (!x.q || !x.t || x.qlen < 1 || x.tlen < 1 || x.h0 < 1 || x.w < 0) return false;
        if ((int64_t)x.h0 + (int64_t)x.qlen * max_mat > SCORE_CAP || x.tlen > 500000) return false;
        if (x.qlen > K1_QLEN_CAP || (opt.force_kernel == 0 && x.qlen >= opt.k2_min_qlen)) return false;
        if (x.q < q0 || x.t < t0) return false;                                   // offsets are taken from the first task
    }
    const ExtTask& l = t[count - 1];
    return (l.q + l.qlen - q0) < (ptrdiff_t)0x7fff0000 && (l.t + l.tlen - t0) < (ptrdiff_t)0x7fff0000;
}

int slot_submit(bsw_ctx* ctx, Slot& s, size_t first, size_t count, int max_mat, const DevParams& dp, int sym,
                const SchedOptions& opt, bool timing, LocalStats* st, bool raw = false, bool allow_dp = false)
{
    const double t0 = now_ms();
    int rc;
    s.raw_mode = raw;
    // upper bounds of the input block: nslots <= count + 4 classes * 31 padding lanes, tiles <= count + 4
    const size_t src_bound = raw ? 0 : source_arena_bound(s.tasks.data(), count) * 4;
    const size_t max_slots = count + 4 * TILE_LANES, max_tiles = count + 4;
    const size_t in_bound = src_bound + max_tiles * sizeof(TileHdr) + max_slots * (sizeof(SlotParam) + sizeof(SlotSrc) + sizeof(uint32_t)) + 64;
    if ((rc = grow_pinned(ctx, &s.h_in, &s.h_in_cap, in_bound))) return rc;
    s.cls.resize(count); s.src.resize(count);
    size_t raw_qbytes = 0, raw_tbytes = 0;
    if (raw) {
        // no staging: the bases stay where the caller has them; a slot's source is a byte offset into the chunk's ranges
        const ExtTask* t = s.tasks.data();
        const uint8_t cl = opt.fast_matrix ? 0 : 1;
        for (size_t i = 0; i < count; ++i) {
            s.cls[i] = cl;
            s.src[i] = SlotSrc{ (uint32_t)(t[i].q - t[0].q), (uint32_t)(t[i].t - t[0].t) };
        }
        raw_qbytes = (size_t)(t[count - 1].q + t[count - 1].qlen - t[0].q);
        raw_tbytes = (size_t)(t[count - 1].t + t[count - 1].tlen - t[0].t);
        s.src_words = 0;
    } else {
        size_t bad = 0; std::string msg;
        rc = pack_tasks(s.tasks.data(), count, max_mat, opt, s.cls.data(), s.src.data(), reinterpret_cast<uint32_t*>(s.h_in),
                        &s.src_words, &bad, &msg);
        if (rc) {
            const size_t colon = msg.find(':');
            set_error(ctx, "task " + std::to_string(first + bad) + (colon == std::string::npos ? "" : msg.substr(colon)));
            return rc;
        }
    }
    const double t1 = now_ms();
    DpGeometry geo{};
    s.dp_mode = allow_dp && build_dp_plan(s.tasks.data(), s.cls.data(), count, opt, &s.plan, &geo);
    if (!s.dp_mode) build_plan(s.tasks.data(), s.cls.data(), s.src.data(), count, opt, &s.plan);
    Plan& P = s.plan;
    const size_t nslots = s.dp_mode ? geo.nslots : P.slots.size();
    s.off_tiles = (s.src_words * 4 + 15) & ~(size_t)15;
    s.off_slots = s.off_tiles + P.tiles.size() * sizeof(TileHdr);
    DpArgs da{};
    if (s.dp_mode) {
        // input block: [source arena | TileHdr[] | task SlotParam[count] | task SlotSrc[count] | task class[count]]
        s.off_ssrc = s.off_slots + count * sizeof(SlotParam);
        s.off_oidx = s.off_ssrc + count * sizeof(SlotSrc);
        s.in_bytes = (s.off_oidx + count + 15) & ~(size_t)15;
        // device arrays the planner writes: slot scalars, slot sources, slot -> task index
        s.dp_off_ssrc = nslots * sizeof(SlotParam);
        s.dp_off_oidx = s.dp_off_ssrc + nslots * sizeof(SlotSrc);
        const size_t off_bin = (s.dp_off_oidx + nslots * sizeof(uint32_t) + 15) & ~(size_t)15;
        const size_t off_tbin = off_bin + (size_t)geo.nmajor * dp_bins_per_major() * sizeof(uint32_t);
        if ((rc = grow_device(ctx, &s.d_plan, &s.d_plan_cap, off_tbin + count * sizeof(uint32_t)))) return rc;
        da.count = (uint32_t)count; da.ntiles = geo.ntiles; da.nmajor = geo.nmajor;
        memcpy(da.major_start, geo.major_start, sizeof(da.major_start)); memcpy(da.major_of, geo.major_of, sizeof(da.major_of));
        da.bins = reinterpret_cast<uint32_t*>(s.d_plan + off_bin); da.task_bin = reinterpret_cast<uint32_t*>(s.d_plan + off_tbin);
        for (int c = 0; c < 2; ++c) { da.class_count[c] = geo.class_count[c]; da.class_pos0[c] = geo.class_pos0[c]; da.class_slot0[c] = geo.class_slot0[c]; da.class_tile0[c] = geo.class_tile0[c]; }
        da.slots = reinterpret_cast<SlotParam*>(s.d_plan); da.slot_src = reinterpret_cast<SlotSrc*>(s.d_plan + s.dp_off_ssrc);
        da.out_index = reinterpret_cast<uint32_t*>(s.d_plan + s.dp_off_oidx);
    } else {
        s.off_ssrc = s.off_slots + nslots * sizeof(SlotParam);
        s.off_oidx = s.off_ssrc + nslots * sizeof(SlotSrc);
        s.in_bytes = s.off_oidx + nslots * sizeof(uint32_t);
    }
    if (s.in_bytes > s.h_in_cap) { set_error(ctx, "internal: input block bound exceeded"); return BSW_ENOMEM; }
    if ((rc = grow_pinned(ctx, &s.h_out, &s.h_out_cap, count))) return rc;
    if ((rc = grow_device(ctx, &s.d_in, &s.d_in_cap, in_bound))) return rc;
    if ((rc = grow_device(ctx, &s.d_arena, &s.d_arena_cap, P.tiled_words))) return rc;
    if ((rc = grow_device(ctx, &s.d_out, &s.d_out_cap, count))) return rc;      // results come back in task order
    if (raw) {
        if ((rc = grow_device(ctx, &s.d_rawq, &s.d_rawq_cap, raw_qbytes + 32))) return rc;       // + the gather's 15-byte over-read
        if ((rc = grow_device(ctx, &s.d_rawt, &s.d_rawt_cap, raw_tbytes + 32))) return rc;
        if ((rc = grow_device(ctx, &s.d_flags, &s.d_flags_cap, nslots))) return rc;
    }
    memcpy(s.h_in + s.off_tiles, P.tiles.data(), P.tiles.size() * sizeof(TileHdr));
    if (s.dp_mode) {
        SlotParam* tp = reinterpret_cast<SlotParam*>(s.h_in + s.off_slots);
        const ExtTask* t = s.tasks.data();
        for (size_t k = 0; k < count; ++k) tp[k] = SlotParam{ t[k].qlen, t[k].tlen, t[k].h0, t[k].w };
        memcpy(s.h_in + s.off_ssrc, s.src.data(), count * sizeof(SlotSrc));
        memcpy(s.h_in + s.off_oidx, s.cls.data(), count);
        da.task_param = reinterpret_cast<const SlotParam*>(s.d_in + s.off_slots);
        da.task_src = reinterpret_cast<const SlotSrc*>(s.d_in + s.off_ssrc);
        da.task_cls = reinterpret_cast<const uint8_t*>(s.d_in + s.off_oidx);
        da.tiles = reinterpret_cast<TileHdr*>(s.d_in + s.off_tiles);
    } else {
        memcpy(s.h_in + s.off_slots, P.slots.data(), nslots * sizeof(SlotParam));
        memcpy(s.h_in + s.off_ssrc, P.slot_src.data(), nslots * sizeof(SlotSrc));
        uint32_t* oi = reinterpret_cast<uint32_t*>(s.h_in + s.off_oidx);
        for (size_t k = 0; k < nslots; ++k) oi[k] = (uint32_t)P.slot_task[k];      // padding lanes (-1) never write
    }
    const double t2 = now_ms();

    // per chunk: 1 H2D, 1 gather, the bucket launches, 1 D2H, 1 event (the per-task cells come back in the records)
    if (timing) CUDA_TRY(ctx, cudaEventRecord(s.ev_in, s.stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(s.d_in, s.h_in, s.in_bytes, cudaMemcpyHostToDevice, s.stream));
    if (raw) {                                           // straight from the caller's registered buffers
        CUDA_TRY(ctx, cudaMemcpyAsync(s.d_rawq, s.tasks[0].q, raw_qbytes, cudaMemcpyHostToDevice, s.stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(s.d_rawt, s.tasks[0].t, raw_tbytes, cudaMemcpyHostToDevice, s.stream));
    }
    if (timing) CUDA_TRY(ctx, cudaEventRecord(s.ev_k0, s.stream));
    if (s.dp_mode) {
        const cudaError_t e = dp_plan_launch(da, s.stream);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "device planner launch");
    }
    if ((rc = enqueue_gather(ctx, s))) return rc;
    if ((rc = enqueue_launches(ctx, s, dp, sym, opt.variant, false, &s.nlaunch, s.d_oidx()))) return rc;
    if (s.dp_mode) s.nlaunch += 3;
    if (timing) CUDA_TRY(ctx, cudaEventRecord(s.ev_k1, s.stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(s.h_out, s.d_out, count * sizeof(SlotResult), cudaMemcpyDeviceToHost, s.stream));
    if (timing) CUDA_TRY(ctx, cudaEventRecord(s.ev_out, s.stream));
    CUDA_TRY(ctx, cudaEventRecord(s.ev_done, s.stream));
    s.timed = timing;
    s.busy = true; s.first = first; s.count = count;
    if (P.n_k1_tiles) ++s.nlaunch;
    st->validate_ms += t1 - t0;
    st->pack_ms += t2 - t1;
    s.trace_ms[0] = t1 - t0; s.trace_ms[1] = t2 - t1; s.trace_ms[2] = now_ms() - t2;
    st->h2d += s.in_bytes + raw_qbytes + raw_tbytes;
    st->d2h += count * sizeof(SlotResult);
    st->launches += s.nlaunch;
    return 0;
}


// ---- lean path: flat batch, bases in registered memory, scheduling on the device ----
// One pass over the caller's offset / h0 / w arrays writes the chunk's tasks straight into the pinned input block
// (24 bytes read, 24 written per task: no ExtTask records, no class bytes, no slot arrays) and keeps the per-bucket
// counts the launch geometry needs; the bases are copied by the DMA engine where they lie; the kernels leave the final
// 24-byte records in task order, so the results need one D2H -- into the caller's array itself when that is registered
// too -- and the host never touches them.  Returns 1 when the chunk is not eligible (the caller then takes the
// general path, which also words the error messages), 0 on success, < 0 on error.
int slot_submit_lean(bsw_ctx* ctx, Slot& s, const FlatSrc& F, size_t first, size_t count, int max_mat, const DevParams& dp, int sym,
                     const SchedOptions& opt, bool timing, LocalStats* st, bsw_result* out, bool out_registered, bool want_cells,
                     bool packed2, std::vector<size_t>* rerun_n)
{
    const double t0 = now_ms();
    if (opt.force_kernel == 2 || count == 0) return 1;
    int rc;
    const size_t max_tiles = count / TILE_LANES + 8;
    const size_t off_param = 0, off_src = count * sizeof(SlotParam);
    const size_t off_tiles = (off_src + count * sizeof(SlotSrc) + 15) & ~(size_t)15;
    const size_t off_arena2 = off_tiles + max_tiles * sizeof(TileHdr);
    size_t in_bound = off_arena2 + 64;
    if (packed2) {
        const int64_t qb = F.qoff[first + count] - F.qoff[first], tb = F.toff[first + count] - F.toff[first];
        if (qb < 0 || tb < 0 || qb >= 0x7fff0000 || tb >= 0x7fff0000) return 1;
        in_bound += (size_t)(qb + tb) / 4 + 8 * count + 128;            // 2 bit per base, every sequence rounded up to a word
    }
    if ((rc = grow_pinned(ctx, &s.h_in, &s.h_in_cap, in_bound))) return rc;
    SlotParam* tp = reinterpret_cast<SlotParam*>(s.h_in + off_param);
    SlotSrc* ts = reinterpret_cast<SlotSrc*>(s.h_in + off_src);
    const int64_t* qoff = F.qoff + first; const int64_t* toff = F.toff + first;
    const int32_t* h0 = F.h0 + first; const int32_t* w = F.w + first;
    const int64_t q0 = qoff[0], tt0 = toff[0];
    const int qcap = opt.force_kernel == 0 ? std::min(K1_QLEN_CAP, opt.k2_min_qlen - 1) : K1_QLEN_CAP;
    const int cl = opt.fast_matrix ? 0 : 1;
    const BandClamp& clamp = *F.clamp;
    DpBuckets bks;
    memset(&bks, 0, sizeof(bks));
    for (size_t k = 0; k < count; ++k) {
        const int64_t ql = qoff[k + 1] - qoff[k], tl = toff[k + 1] - toff[k];
        const int64_t qo = qoff[k] - q0, to = toff[k] - tt0;
        if (ql < 1 || tl < 1 || ql > qcap || tl > 500000 || h0[k] < 1 || w[k] < 0 || qo < 0 || to < 0 ||
            qo >= 0x7fff0000 || to >= 0x7fff0000 || (int64_t)h0[k] + ql * max_mat > SCORE_CAP) return 1;
        const int wc = clamp((int)ql, w[k]);
        tp[k] = SlotParam{ (int32_t)ql, (int32_t)tl, h0[k], wc };
        ts[k] = SlotSrc{ (uint32_t)qo, (uint32_t)to };
        bks.add(cl, (int)ql, (int)tl, wc);
    }
    size_t raw_qbytes = (size_t)(qoff[count] - q0), raw_tbytes = (size_t)(toff[count] - tt0);
    if (raw_qbytes >= 0x7fff0000u || raw_tbytes >= 0x7fff0000u) return 1;
    size_t arena2_bytes = 0;
    if (packed2) {
        std::vector<uint32_t> n_tasks;
        const int64_t words = pack2_flat(F.qbuf, qoff, F.tbuf, toff, count, reinterpret_cast<uint32_t*>(s.h_in + off_arena2), ts, &n_tasks);
        if (words < 0) { set_error(ctx, "task " + std::to_string(first + (size_t)(-1 - words)) + ": invalid (base code > 4)"); return BSW_EINVAL; }
        for (uint32_t k : n_tasks) rerun_n->push_back(first + k);            // an N has no 2-bit code: those tasks are rerun (4 bit, matrix lookup)
        arena2_bytes = (size_t)words * 4 + 16;
        raw_qbytes = raw_tbytes = 0;
    }
    const double t1 = now_ms();
    DpGeometry geo{};
    if (!dp_geometry(bks, count, opt, &s.plan, &geo)) return 1;
    Plan& P = s.plan;
    if (P.tiles.size() > max_tiles) { set_error(ctx, "internal: tile bound exceeded"); return BSW_ENOMEM; }
    const size_t nslots = geo.nslots;
    s.raw_mode = !packed2; s.dp_mode = true; s.lean = true; s.lean_direct = out_registered; s.lean_cells = want_cells;
    s.off_arena2 = packed2 ? off_arena2 : 0;
    s.src_words = 0;
    s.off_tiles = off_tiles; s.off_slots = off_param; s.off_ssrc = off_src; s.off_oidx = 0;
    s.in_bytes = packed2 ? off_arena2 + arena2_bytes : off_tiles + P.tiles.size() * sizeof(TileHdr);
    memcpy(s.h_in + off_tiles, P.tiles.data(), P.tiles.size() * sizeof(TileHdr));
    s.dp_off_ssrc = nslots * sizeof(SlotParam);
    s.dp_off_oidx = s.dp_off_ssrc + nslots * sizeof(SlotSrc);
    const size_t off_bin = (s.dp_off_oidx + nslots * sizeof(uint32_t) + 15) & ~(size_t)15;
    const size_t off_tbin = off_bin + (size_t)geo.nmajor * dp_bins_per_major() * sizeof(uint32_t);
    if ((rc = grow_device(ctx, &s.d_plan, &s.d_plan_cap, off_tbin + count * sizeof(uint32_t)))) return rc;
    if ((rc = grow_device(ctx, &s.d_in, &s.d_in_cap, in_bound))) return rc;
    if ((rc = grow_device(ctx, &s.d_arena, &s.d_arena_cap, P.tiled_words))) return rc;
    if (!packed2) {
        if ((rc = grow_device(ctx, &s.d_rawq, &s.d_rawq_cap, raw_qbytes + 32))) return rc;
        if ((rc = grow_device(ctx, &s.d_rawt, &s.d_rawt_cap, raw_tbytes + 32))) return rc;
        if ((rc = grow_device(ctx, &s.d_flags, &s.d_flags_cap, nslots))) return rc;
    }
    if ((rc = grow_device(ctx, &s.d_out24, &s.d_out24_cap, count * 6))) return rc;
    if (!out_registered && (rc = grow_pinned(ctx, &s.h_out24, &s.h_out24_cap, count * 6))) return rc;
    if (want_cells) {
        if ((rc = grow_device(ctx, &s.d_cellsv, &s.d_cellsv_cap, count))) return rc;
        if ((rc = grow_pinned(ctx, &s.h_cellsv, &s.h_cellsv_cap, count))) return rc;
    }
    if (!s.d_flaglist) {
        CUDA_TRY(ctx, cudaMalloc((void**)&s.d_flaglist, (LEAN_FLAGS + 1) * sizeof(uint32_t)));
        CUDA_TRY(ctx, cudaHostAlloc((void**)&s.h_flaglist, (LEAN_FLAGS + 1) * sizeof(uint32_t), cudaHostAllocDefault));
    }
    DpArgs da{};
    da.count = (uint32_t)count; da.ntiles = geo.ntiles; da.nmajor = geo.nmajor;
    memcpy(da.major_start, geo.major_start, sizeof(da.major_start)); memcpy(da.major_of, geo.major_of, sizeof(da.major_of));
    for (int c = 0; c < 2; ++c) { da.class_count[c] = geo.class_count[c]; da.class_pos0[c] = geo.class_pos0[c]; da.class_slot0[c] = geo.class_slot0[c]; da.class_tile0[c] = geo.class_tile0[c]; }
    da.bins = reinterpret_cast<uint32_t*>(s.d_plan + off_bin); da.task_bin = reinterpret_cast<uint32_t*>(s.d_plan + off_tbin);
    da.slots = reinterpret_cast<SlotParam*>(s.d_plan); da.slot_src = reinterpret_cast<SlotSrc*>(s.d_plan + s.dp_off_ssrc);
    da.out_index = reinterpret_cast<uint32_t*>(s.d_plan + s.dp_off_oidx);
    da.task_param = reinterpret_cast<const SlotParam*>(s.d_in + off_param);
    da.task_src = reinterpret_cast<const SlotSrc*>(s.d_in + off_src);
    da.task_cls = nullptr; da.const_cls = (uint32_t)cl;
    da.tiles = reinterpret_cast<TileHdr*>(s.d_in + off_tiles);
    const double t2 = now_ms();

    if (timing) CUDA_TRY(ctx, cudaEventRecord(s.ev_in, s.stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(s.d_in, s.h_in, s.in_bytes, cudaMemcpyHostToDevice, s.stream));
    if (!packed2) {
        CUDA_TRY(ctx, cudaMemcpyAsync(s.d_rawq, F.qbuf + q0, raw_qbytes, cudaMemcpyHostToDevice, s.stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(s.d_rawt, F.tbuf + tt0, raw_tbytes, cudaMemcpyHostToDevice, s.stream));
    }
    CUDA_TRY(ctx, cudaMemsetAsync(s.d_flaglist, 0, sizeof(uint32_t), s.stream));
    CUDA_TRY(ctx, cudaMemsetAsync(s.d_cells, 0, sizeof(unsigned long long), s.stream));
    if (timing) CUDA_TRY(ctx, cudaEventRecord(s.ev_k0, s.stream));
    {
        const cudaError_t e = dp_plan_launch(da, s.stream);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "device planner launch");
    }
    if ((rc = enqueue_gather(ctx, s))) return rc;
    if ((rc = enqueue_launches(ctx, s, dp, sym, opt.variant, true, &s.nlaunch, s.d_oidx()))) return rc;
    if (timing) CUDA_TRY(ctx, cudaEventRecord(s.ev_k1, s.stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(out_registered ? (void*)(out + first) : (void*)s.h_out24, s.d_out24, count * sizeof(bsw_result),
                                  cudaMemcpyDeviceToHost, s.stream));
    if (want_cells) CUDA_TRY(ctx, cudaMemcpyAsync(s.h_cellsv, s.d_cellsv, count * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(s.h_flaglist, s.d_flaglist, (LEAN_FLAGS + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(s.h_cells, s.d_cells, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
    if (timing) CUDA_TRY(ctx, cudaEventRecord(s.ev_out, s.stream));
    CUDA_TRY(ctx, cudaEventRecord(s.ev_done, s.stream));
    s.timed = timing;
    s.busy = true; s.first = first; s.count = count;
    s.nlaunch += 4;                                         // 3 planner kernels + the gather
    st->validate_ms += t1 - t0;
    st->pack_ms += t2 - t1;
    s.trace_ms[0] = t1 - t0; s.trace_ms[1] = t2 - t1; s.trace_ms[2] = now_ms() - t2;
    st->h2d += s.in_bytes + raw_qbytes + raw_tbytes;
    st->d2h += count * sizeof(bsw_result) + (want_cells ? count * sizeof(uint32_t) : 0) + (LEAN_FLAGS + 1) * sizeof(uint32_t);
    st->launches += s.nlaunch;
    return 0;
}

int slot_collect_lean(bsw_ctx* ctx, Slot& s, bsw_result* out, uint32_t* cells, LocalStats* st, std::vector<size_t>* rerun_n)
{
    CUDA_TRY(ctx, cudaEventSynchronize(s.ev_done));
    s.busy = false; s.lean = false;
    float ms = 0.f;
    if (s.timed) CUDA_TRY(ctx, cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1));
    if (s.timed && ctx->trace_ref) {
        float a = 0, b = 0, c = 0, d = 0;
        cudaEventElapsedTime(&a, ctx->trace_ref, s.ev_in); cudaEventElapsedTime(&b, ctx->trace_ref, s.ev_k0);
        cudaEventElapsedTime(&c, ctx->trace_ref, s.ev_k1); cudaEventElapsedTime(&d, ctx->trace_ref, s.ev_out);
        fprintf(stderr, "gpu chunk %zu+%zu: h2d %.3f..%.3f kernels ..%.3f d2h ..%.3f ms\n", s.first, s.count, a, b, c, d);
    }
    const size_t first = s.first, count = s.count;
    if (!s.lean_direct) memcpy(out + first, s.h_out24, count * sizeof(bsw_result));
    if (cells && s.lean_cells) memcpy(cells + first, s.h_cellsv, count * sizeof(uint32_t));
    uint32_t nflag = s.h_flaglist[0];
    std::vector<uint32_t> more;
    const uint32_t* list = s.h_flaglist + 1;
    if (nflag > LEAN_FLAGS) {
        // more flagged tasks than the first copy holds (a batch full of N): the kernels stopped recording at the cap,
        // so every task of the chunk is rerun on the staged path -- correct, just slower, and rare
        for (size_t t = 0; t < count; ++t) rerun_n->push_back(first + t);
        nflag = 0;
    }
    for (uint32_t k = 0; k < nflag; ++k) {
        const uint32_t status = list[k] >> 28; const size_t t = first + (list[k] & 0x0fffffffu);
        if (status == STATUS_HAS_N) rerun_n->push_back(t);
        else if (status == STATUS_BAD_CODE) { set_error(ctx, "task " + std::to_string(t) + ": invalid (base code > 4)"); return BSW_EINVAL; }
        else { set_error(ctx, "a kernel reported a non-OK task status"); return BSW_ECUDA; }
    }
    st->tasks += count; st->cells += *s.h_cells; st->kernel_ms += ms;
    return 0;
}

// Wait for the slot's chunk and scatter its results to out[first + task].
int slot_collect(bsw_ctx* ctx, Slot& s, bsw_result* out, uint32_t* cells, LocalStats* st, std::vector<size_t>* rerun_n)
{
    if (!s.busy) return 0;
    if (s.lean) return slot_collect_lean(ctx, s, out, cells, st, rerun_n);
    CUDA_TRY(ctx, cudaEventSynchronize(s.ev_done));
    s.busy = false;
    float ms = 0.f;
    if (s.timed) CUDA_TRY(ctx, cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1));
    if (s.timed && ctx->trace_ref) {
        float a = 0, b = 0, c = 0, d = 0;
        cudaEventElapsedTime(&a, ctx->trace_ref, s.ev_in); cudaEventElapsedTime(&b, ctx->trace_ref, s.ev_k0);
        cudaEventElapsedTime(&c, ctx->trace_ref, s.ev_k1); cudaEventElapsedTime(&d, ctx->trace_ref, s.ev_out);
        fprintf(stderr, "gpu chunk %zu+%zu: h2d %.3f..%.3f kernels ..%.3f d2h ..%.3f ms\n", s.first, s.count, a, b, c, d);
    }
    const size_t first = s.first, count = s.count;
    const SlotResult* h_out = s.h_out;
    int bad = 0;
    uint64_t cell_sum = 0;
    // the kernels wrote every record at its task's index: a sequential pass
    for (size_t t = 0; t < count; ++t) {
        const SlotResult& r = h_out[t];
        if (r.status == STATUS_HAS_N) { rerun_n->push_back(first + t); continue; }         // raw mode: rerun with matrix lookup
        if (r.status == STATUS_BAD_CODE) {
            set_error(ctx, "task " + std::to_string(first + t) + ": invalid (base code > 4)");
            return BSW_EINVAL;
        }
        if (r.status != STATUS_OK) bad = 1;
        cell_sum += (uint32_t)r.cells;
        bsw_result& o = out[first + t];
        o.score = r.score; o.qle = r.qle; o.tle = r.tle; o.gtle = r.gtle; o.gscore = r.gscore; o.max_off = r.max_off;
        if (cells) cells[first + t] = (uint32_t)r.cells;
    }
    st->tasks += s.count; st->cells += cell_sum; st->kernel_ms += ms;
    if (bad) { set_error(ctx, "a kernel reported a non-OK task status"); return BSW_ECUDA; }
    return 0;
}

// Runs fn(0..n-1) on n threads (the caller is worker 0).  Creating a thread costs ~17 us, so the second half of the
// workers is created by the first spawned thread while the caller creates the first half: the last worker starts after
// ~n/2 creations instead of n (0.3 -> 0.15 ms for 16 workers, which matters for 100 k task batches of ~1 ms).
template <class F>
void run_workers(size_t n, F& fn)
{
    if (n <= 1) { fn(0); return; }
    const size_t half = n >= 6 ? (n + 1) / 2 : n;        // workers [half, n) belong to the helper
    std::thread helper;
    if (half < n)
        helper = std::thread([&fn, half, n]() {
            std::vector<std::thread> th;
            for (size_t k = half + 1; k < n; ++k) th.emplace_back([&fn, k]() { fn(k); });
            fn(half);
            for (auto& t : th) t.join();
        });
    std::vector<std::thread> th;
    for (size_t k = 1; k < half; ++k) th.emplace_back([&fn, k]() { fn(k); });
    fn(0);
    for (auto& t : th) t.join();
    if (helper.joinable()) helper.join();
}

void adopt_worker_error(bsw_ctx* ctx)
{
    std::lock_guard<std::mutex> g(ctx->err_mu);
    tl_last_error = ctx->last_error;
}

Worker* get_worker(bsw_ctx* ctx, bsw_ctx::CallState& cs, size_t k)
{
    while (cs.workers.size() <= k) {
        std::unique_ptr<Worker> w(new Worker());
        w->dev = (int)(cs.workers.size() % ctx->devs.size());        // workers are dealt round-robin over the devices
        cs.workers.push_back(std::move(w));
    }
    return cs.workers[k].get();
}

// RAII: a CallState for the duration of one batch call
struct CallLease {
    bsw_ctx* ctx; bsw_ctx::CallState* cs;
    explicit CallLease(bsw_ctx* c) : ctx(c), cs(nullptr)
    {
        ctx->calls_live.fetch_add(1);
        std::lock_guard<std::mutex> g(ctx->pool_mu);
        if (!ctx->call_free.empty()) { cs = ctx->call_free.back().release(); ctx->call_free.pop_back(); }
        else cs = new bsw_ctx::CallState();
    }
    ~CallLease()
    {
        ctx->calls_live.fetch_sub(1);
        std::lock_guard<std::mutex> g(ctx->pool_mu);
        ctx->call_free.emplace_back(cs);
    }
};

// The engine behind every batch entry point.  The batch is cut into chunks; `nworkers` host threads each run a
// pipeline (fill -> validate -> plan -> pack -> H2D -> kernels -> D2H -> scatter) over the chunks they pull from a
// shared counter, alternating between their two stream slots.  Workers are bound round-robin to the devices, so a
// multi-GPU context balances dynamically -- the GPU analogue of task_parse handing the next task to the first PE
// with room (sw_pe_array_task_parse.v:1600-1650).  No collective: results land in out[task].
int run_extensions_locked(bsw_ctx* ctx, bsw_ctx::CallState& cs, const bsw_params* params, const TaskSource& src, size_t n, bsw_result* out,
                          uint32_t* cells, std::vector<size_t>* rerun_n_out = nullptr)
{
    const double w0 = now_ms();
    DevParams dp; int sym = 0, max_mat = 0; bool fast_ok = false;
    int rc = make_dev_params(ctx, params, &dp, &sym, &fast_ok, &max_mat);
    if (rc) return rc;
    SchedOptions opt = ctx->opt;
    opt.fast_matrix = fast_ok;
    if (opt.host_threads <= 0) opt.host_threads = default_host_threads();

    // chunk size: chunk_tasks for short reads, fewer tasks per chunk when they are long (about 6 MB of bases per chunk),
    // so that a batch of long tasks still spreads over all the host workers
    // Raw mode (bases in registered host memory, no staging pass) trades host work for PCIe bytes: measured on
    // 1 M x 150 bp it wins below ~10 host threads per GPU (4 threads: 7.5 vs 10.6 ms, 8: 6.6 vs 7.1) and loses above
    // (16 threads: 6.5 vs 5.8 ms), so "auto" takes it only when the host is the scarce side.  Its copies are per
    // buffer, so it runs on chunks twice the usual size.
    const size_t ndev = ctx->devs.size();
    // Flat batches planned on the device take the lean path: 2 = the host packs the bases 2 bit each (any memory), 1 = the
    // DMA engine copies them one byte each from registered buffers (no host pass over the bases at all).  Packing costs
    // ~8 ns per task and thread and cuts the host->device bytes 2.5x, so "auto" packs unless a GPU has a single host thread.
    // "auto": raw when the buffers are registered and the GPU has fewer than 12 host threads to itself, packed otherwise.  Measured on an 8-GPU box (32 host cores, 4 per
    // GPU): raw 11.3 ms per step (GPUs 0-3 sit behind a slower root complex: 20 GB/s each when all eight copy), packed
    // 17.6 ms (the cores read 8 x 190 MB per step: the host memory system is the limit), a per-chunk switch on "are the
    // previous chunk's copies still in flight" 20.1 ms (it packs about half of the chunks).  With 16 cores for one GPU
    // packed is the faster one by 3-10 % (5.5 vs 5.7-6.0 ms per 1 M tasks).
    int lean_mode = 0;
    if (src.flat && ctx->device_plan && rerun_n_out != nullptr) {
        if (ctx->raw_inputs == 3) lean_mode = 2;
        else if (ctx->raw_inputs == 1) lean_mode = src.raw ? 1 : 0;
        else if (ctx->raw_inputs == 2) lean_mode = (src.raw && (size_t)opt.host_threads < 12 * ndev) ? 1 : 2;
    }
    const bool allow_raw = lean_mode == 1 || (lean_mode == 0 && src.raw && rerun_n_out != nullptr &&
                           (ctx->raw_inputs == 1 || (ctx->raw_inputs == 2 && (size_t)opt.host_threads <= 10 * ndev)));
    size_t chunk = std::max<size_t>(32, allow_raw ? 2 * ctx->chunk_tasks : ctx->chunk_tasks);
    {
        const size_t probe = std::min<size_t>(n, 512);
        std::vector<ExtTask> pv(probe);
        src.fill(src.self, 0, probe, pv.data());
        uint64_t bases = 0;
        for (const ExtTask& t : pv) bases += (uint64_t)std::max(t.qlen, 0) + (uint64_t)std::max(t.tlen, 0);
        const size_t mean = (size_t)(bases / probe) + 1;
        const size_t by_bytes = std::max<size_t>(64, (size_t)(6u << 20) / mean);
        chunk = std::min(chunk, by_bytes);
        // a batch smaller than threads x chunk: still one chunk per worker (not below 2048 tasks: launch overheads)
        const size_t per_worker = (n + (size_t)opt.host_threads - 1) / (size_t)opt.host_threads;
        chunk = std::min(chunk, std::max<size_t>(2048, (per_worker + 31) & ~(size_t)31));
    }
    // The lean path leaves the host ~1 us per 1000 tasks of work, so what matters is the pipeline: every worker should own
    // several chunks, or all the copies are issued at once and nothing overlaps (32 host threads, one 31 k task chunk
    // each: 8.0 ms per 1 M tasks against 5.7 ms with 16 threads).  At most 8 workers per GPU, at least ~4 chunks each.
    // calls in flight share the host: four concurrent calls with 32 threads each on a 32-core box just fight
    const size_t live = (size_t)std::max(1, ctx->calls_live.load());
    size_t max_workers = std::max<size_t>(ndev, (size_t)opt.host_threads / live);
    if (lean_mode) {
        max_workers = std::min<size_t>(max_workers, (lean_mode >= 2 ? 16 : 8) * ndev);
        const size_t share = n / (max_workers * 4);
        chunk = std::max<size_t>(std::min(chunk, std::max<size_t>(4096, (share + 31) & ~(size_t)31)), 32);
    }
    const size_t nchunks = (n + chunk - 1) / chunk;
    size_t nworkers = std::min<size_t>(max_workers, nchunks);
    if (nworkers < ndev && nchunks >= ndev) nworkers = ndev;
    if (nworkers < 1) nworkers = 1;
    for (size_t k = 0; k < nworkers; ++k) get_worker(ctx, cs, k);

    // Fixed-size chunks pulled from a shared cursor.  Measured alternatives on 1 M x 150 bp, all slower: chunks that
    // shrink towards the end of the batch (shorter un-overlapped tail, but more launches and copies: 6.5 -> 7.0-9 ms)
    // and a ramp-up of small first chunks (GPU starts earlier: 6.5 -> 6.9 ms; re-measured after the chunk sort key:
    // no difference, 5.8-6.0 ms either way; once more on the lean path, every worker's first chunk a quarter or an eighth
    // of the size: 5.4-5.6 ms either way on 1 M tasks, 1.27 against 1.17 ms on 100 k).  Also without effect: non-temporal
    // stores into the pinned block.
    std::atomic<size_t> cursor(0);
    auto grab = [&](size_t* first, size_t* count) -> bool {
        const size_t cur = cursor.fetch_add(chunk, std::memory_order_relaxed);
        if (cur >= n) return false;
        *first = cur; *count = std::min(chunk, n - cur);
        return true;
    };
    std::atomic<int> first_err(0);
    std::mutex stat_mu;
    LocalStats total;
    std::vector<size_t> rerun_all;

    auto worker_main = [&](size_t k) {
        Worker& W = *cs.workers[k];
        LocalStats st;
        std::vector<size_t> rrn;
        int r = 0;
        if (cudaSetDevice(ctx->devs[(size_t)W.dev].id) != cudaSuccess) { r = BSW_ECUDA; set_error(ctx, "cudaSetDevice failed"); }
        if (!r && W.slots.size() < (size_t)ctx->slots_per_worker) {
            const size_t have = W.slots.size();
            W.slots.resize((size_t)ctx->slots_per_worker);
            for (size_t q = have; q < W.slots.size() && !r; ++q) r = slot_init(ctx, W.slots[q]);
        }
        const size_t nslot = (size_t)ctx->slots_per_worker;
        size_t cur = 0;
        const bool trace = getenv("BSW_TRACE") != nullptr;
        std::string tr;
        auto T = [&]() { return now_ms() - w0; };
        if (trace) tr += "w" + std::to_string(k) + " start " + std::to_string(T()) + "\n";
        while (!r && !first_err.load(std::memory_order_relaxed)) {
            size_t first = 0, count = 0;
            if (!grab(&first, &count)) break;
            Slot& s = W.slots[cur];
            cur = (cur + 1) % nslot;
            const double c0 = T();
            if ((r = slot_collect(ctx, s, out, cells, &st, &rrn))) break;
            if (trace) tr += "w" + std::to_string(k) + " chunk " + std::to_string(first) + "+" + std::to_string(count) + " collect_prev " + std::to_string(c0) + ".." + std::to_string(T());
            if (lean_mode) {
                const bool packed2 = lean_mode == 2;
                r = slot_submit_lean(ctx, s, *src.flat, first, count, max_mat, dp, sym, opt, ctx->kernel_timing, &st, out, src.out_registered, cells != nullptr,
                                     packed2, &rrn);
                if (r == 0) { if (packed2) ++st.packed_chunks; else ++st.raw_chunks; }
                if (r <= 0) {
                    if (trace) tr += " lean submitted " + std::to_string(T()) + " (pass " + std::to_string(s.trace_ms[0]) + " geometry " + std::to_string(s.trace_ms[1]) + " api " + std::to_string(s.trace_ms[2]) + ")\n";
                    continue;
                }
                r = 0;                                   // not eligible: the general path below
            }
            const double v0 = now_ms();
            s.tasks.resize(count);
            src.fill(src.self, first, count, s.tasks.data());
            st.validate_ms += now_ms() - v0;
            const double f1 = T();
            const bool raw = allow_raw && raw_chunk_eligible(s.tasks.data(), count, max_mat, opt);
            s.lean = false;
            r = slot_submit(ctx, s, first, count, max_mat, dp, sym, opt, ctx->kernel_timing, &st, raw, ctx->device_plan);
            if (trace) tr += " fill.." + std::to_string(f1) + " submitted " + std::to_string(T()) + " (pack " + std::to_string(s.trace_ms[0]) +
                             " plan " + std::to_string(s.trace_ms[1]) + " api " + std::to_string(s.trace_ms[2]) + ")\n";
        }
        if (trace) tr += "w" + std::to_string(k) + " drain " + std::to_string(T());
        for (size_t q = 0; q < W.slots.size(); ++q) {                                    // oldest chunk first
            Slot& s = W.slots[(cur + q) % W.slots.size()];
            if (r) { if (s.stream) cudaStreamSynchronize(s.stream); s.busy = false; }    // leave the device quiescent
            else r = slot_collect(ctx, s, out, cells, &st, &rrn);
        }
        if (r) { int expect = 0; first_err.compare_exchange_strong(expect, r); }
        if (trace) { tr += " done " + std::to_string(T()) + "\n"; fputs(tr.c_str(), stderr); }
        std::lock_guard<std::mutex> g(stat_mu);
        rerun_all.insert(rerun_all.end(), rrn.begin(), rrn.end());
        total.pack_ms += st.pack_ms; total.validate_ms += st.validate_ms; total.kernel_ms += st.kernel_ms;
        total.h2d += st.h2d; total.d2h += st.d2h; total.launches += st.launches; total.tasks += st.tasks; total.cells += st.cells;
    };
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    const bool gpu_trace = getenv("BSW_TRACE") != nullptr && ctx->kernel_timing;
    if (gpu_trace) {
        cudaSetDevice(ctx->devs[0].id);
        if (!ctx->trace_ref) cudaEventCreate(&ctx->trace_ref);
        cudaEventRecord(ctx->trace_ref, ctx->devs[0].aux.stream);
        fprintf(stderr, "gpu timeline origin at host %.3f ms\n", now_ms() - w0);
    }
    run_workers(nworkers, worker_main);
    cudaSetDevice(prev_dev);
    {
        std::lock_guard<std::mutex> g(ctx->err_mu);
        bsw_stats& S = ctx->stats;
        S.tasks += total.tasks; S.cells_band += total.cells; S.kernel_launches += total.launches;
        S.h2d_bytes += total.h2d; S.d2h_bytes += total.d2h; S.kernel_ms += total.kernel_ms;
        S.pack_ms += (total.pack_ms + total.validate_ms) / (double)nworkers;     // average per worker = wall share
        S.wall_ms += now_ms() - w0;
        S.tasks -= rerun_all.size();                                // counted again by the rerun
    }
    if (rerun_n_out) rerun_n_out->swap(rerun_all);
    if (first_err.load()) adopt_worker_error(ctx);
    return first_err.load();
}

// Gathers a subset of a task source (a rerun list).
struct SubsetSrc { const TaskSource* base; const size_t* idx; };
void fill_subset(const void* self, size_t first, size_t count, ExtTask* out)
{
    const SubsetSrc& S = *static_cast<const SubsetSrc*>(self);
    for (size_t k = 0; k < count; ++k) S.base->fill(S.base->self, S.idx[first + k], 1, out + k);
}

// 0: the 16-bit kernels take the task; 1: only K5 can; BSW_EINVAL / BSW_ERANGE: nobody can
int wide_class(const ExtTask& t, int max_mat, const DevParams& dp)
{
    if (!t.q || !t.t || t.qlen < 1 || t.tlen < 1 || t.h0 < 1 || t.w < 0) return BSW_EINVAL;
    if ((int64_t)t.h0 + (int64_t)t.qlen * max_mat <= SCORE_CAP && t.qlen <= K2_QLEN_CAP && t.tlen <= 500000) return 0;
    const int64_t e = std::max(dp.e_del, dp.e_ins);
    if (t.qlen > WIDE_QLEN_CAP || t.tlen > WIDE_TLEN_CAP || (int64_t)t.h0 + (int64_t)t.qlen * ((int64_t)max_mat + dp.e_ins) > WIDE_SCORE_CAP ||
        (int64_t)std::max(t.qlen, t.tlen) * e > WIDE_SCORE_CAP) return BSW_ERANGE;
    return 1;
}

// The tasks idx[0..m) (or tasks 0..m when idx is null) on K5: chunks bounded by bases and row workspace, one stream on
// the context's first device.  A correctness path for rare tasks, not a pipeline.
int run_wide(bsw_ctx* ctx, const bsw_params* params, const TaskSource& src, const size_t* idx, size_t m, bsw_result* out, uint32_t* cells)
{
    DevParams dp; int sym = 0, max_mat = 0; bool fast_ok = false;
    int rc = make_dev_params(ctx, params, &dp, &sym, &fast_ok, &max_mat);
    if (rc) return rc;
    int prev = 0, sms = 148;
    cudaGetDevice(&prev);
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[0].id));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->devs[0].id);
    cudaStream_t st = nullptr;
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    auto fail = [&](int code) { cudaStreamDestroy(st); cudaSetDevice(prev); return code; };
    std::vector<WideTask> wt;
    std::vector<uint8_t> qb, tb;
    std::vector<SlotResult> res;
    size_t first = 0;
    uint64_t launches = 0, ntask = 0, ncells = 0;
    while (first < m) {
        wt.clear(); qb.clear(); tb.clear();
        uint64_t rows = 0;
        size_t count = 0;
        while (first + count < m && count < 65536 && qb.size() + tb.size() < ((size_t)256 << 20) && rows < ((uint64_t)64 << 20)) {
            const size_t task = idx ? idx[first + count] : first + count;
            ExtTask t;
            src.fill(src.self, task, 1, &t);
            const int c = wide_class(t, max_mat, dp);
            if (c < 0) {
                set_error(ctx, "task " + std::to_string(task) + ": qlen=" + std::to_string(t.qlen) + " tlen=" + std::to_string(t.tlen) + " h0=" + std::to_string(t.h0) +
                               " w=" + std::to_string(t.w) + (c == BSW_ERANGE ? " outside the numeric envelope (32-bit row state / length caps)"
                                                                             : " invalid (null pointer, length < 1, h0 < 1 or negative band)"));
                return fail(c);
            }
            for (int k = 0; k < t.qlen; ++k) if (t.q[k] > 4) { set_error(ctx, "task " + std::to_string(task) + ": invalid (base code > 4)"); return fail(BSW_EINVAL); }
            for (int k = 0; k < t.tlen; ++k) if (t.t[k] > 4) { set_error(ctx, "task " + std::to_string(task) + ": invalid (base code > 4)"); return fail(BSW_EINVAL); }
            wt.push_back(WideTask{ rows, (uint32_t)qb.size(), (uint32_t)tb.size(), t.qlen, t.tlen, t.h0, t.w });
            rows += 2 * ((uint64_t)t.qlen + 2);
            qb.insert(qb.end(), t.q, t.q + t.qlen); tb.insert(tb.end(), t.t, t.t + t.tlen);
            ++count;
        }
        void *d_tasks = nullptr, *d_q = nullptr, *d_t = nullptr, *d_rows = nullptr, *d_out = nullptr;
        auto release = [&]() { for (void* p : { d_tasks, d_q, d_t, d_rows, d_out }) if (p) cudaFree(p); };
#define W_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { release(); int c__ = cuda_fail(ctx, e__, #call); return fail(c__); } } while (0)
        W_TRY(cudaMalloc(&d_tasks, count * sizeof(WideTask)));
        W_TRY(cudaMalloc(&d_q, qb.size() + 16)); W_TRY(cudaMalloc(&d_t, tb.size() + 16));
        W_TRY(cudaMalloc(&d_rows, (size_t)rows * sizeof(int32_t)));
        W_TRY(cudaMalloc(&d_out, count * sizeof(SlotResult)));
        W_TRY(cudaMemcpyAsync(d_tasks, wt.data(), count * sizeof(WideTask), cudaMemcpyHostToDevice, st));
        W_TRY(cudaMemcpyAsync(d_q, qb.data(), qb.size(), cudaMemcpyHostToDevice, st));
        W_TRY(cudaMemcpyAsync(d_t, tb.data(), tb.size(), cudaMemcpyHostToDevice, st));
        WideArgs a{};
        a.tasks = static_cast<const WideTask*>(d_tasks); a.qbuf = static_cast<const uint8_t*>(d_q); a.tbuf = static_cast<const uint8_t*>(d_t);
        a.rows = static_cast<int32_t*>(d_rows); a.out = static_cast<SlotResult*>(d_out); a.ntasks = (uint32_t)count; a.p = dp;
        W_TRY(k5_launch(a, ctx->opt.variant, sms, st));
        res.resize(count);
        W_TRY(cudaMemcpyAsync(res.data(), d_out, count * sizeof(SlotResult), cudaMemcpyDeviceToHost, st));
        W_TRY(cudaStreamSynchronize(st));
#undef W_TRY
        release();
        for (size_t k = 0; k < count; ++k) {
            const size_t task = idx ? idx[first + k] : first + k;
            const SlotResult& r = res[k];
            bsw_result& o = out[task];
            o.score = r.score; o.qle = r.qle; o.tle = r.tle; o.gtle = r.gtle; o.gscore = r.gscore; o.max_off = r.max_off;
            if (cells) cells[task] = (uint32_t)r.cells;
            ncells += (uint32_t)r.cells;
        }
        ++launches; ntask += count;
        first += count;
    }
    cudaStreamDestroy(st);
    cudaSetDevice(prev);
    std::lock_guard<std::mutex> g(ctx->err_mu);
    ctx->stats.kernel_launches += launches; ctx->stats.tasks += ntask; ctx->stats.cells_band += ncells;
    return BSW_OK;
}

// One pass over a source plus its rerun: raw-mode tasks that hold an N go through the staged path (which classifies them
// for the matrix-lookup kernel) -- whole tasks, from scratch, results scattered over the first pass
int run_with_reruns(bsw_ctx* ctx, bsw_ctx::CallState& cs, const bsw_params* params, const TaskSource& src, size_t n, bsw_result* out, uint32_t* cells)
{
    std::vector<size_t> rerun_n;
    int rc = run_extensions_locked(ctx, cs, params, src, n, out, cells, &rerun_n);
    if (rc || rerun_n.empty()) return rc;
    std::sort(rerun_n.begin(), rerun_n.end());
    const SubsetSrc sub{ &src, rerun_n.data() };
    std::vector<bsw_result> r2(rerun_n.size());
    std::vector<uint32_t> c2(cells ? rerun_n.size() : 0);
    rc = run_extensions_locked(ctx, cs, params, TaskSource{ &sub, fill_subset }, rerun_n.size(), r2.data(), cells ? c2.data() : nullptr, nullptr);
    if (rc) return rc;
    for (size_t k = 0; k < rerun_n.size(); ++k) { out[rerun_n[k]] = r2[k]; if (cells) cells[rerun_n[k]] = c2[k]; }
    return BSW_OK;
}

int run_extensions(bsw_ctx* ctx, const bsw_params* params, const TaskSource& src, size_t n, bsw_result* out, uint32_t* cells)
{
    if (!ctx || !out) { set_error(ctx, "null argument"); return BSW_EINVAL; }
    if (n == 0) return BSW_OK;
    if (ctx->wide == 2) return run_wide(ctx, params, src, nullptr, n, out, cells);
    CallLease lease(ctx);
    bsw_ctx::CallState& cs = *lease.cs;
    int rc = run_with_reruns(ctx, cs, params, src, n, out, cells);
    if (rc != BSW_ERANGE || ctx->wide == 0) return rc;
    // Some task is outside the 16-bit envelope (the batch was refused as a whole; rare).  Split it: the tasks K1 / K2 can
    // take run again as a subset, the others on K5.  A task nobody can take fails the call here with its own message.
    DevParams dp; int sym = 0, max_mat = 0; bool fast_ok = false;
    if ((rc = make_dev_params(ctx, params, &dp, &sym, &fast_ok, &max_mat))) return rc;
    std::vector<size_t> narrow, wide;
    {
        std::vector<ExtTask> buf(4096);
        for (size_t first = 0; first < n; first += buf.size()) {
            const size_t count = std::min(buf.size(), n - first);
            src.fill(src.self, first, count, buf.data());
            for (size_t k = 0; k < count; ++k) {
                const int c = wide_class(buf[k], max_mat, dp);
                if (c == 1) wide.push_back(first + k); else narrow.push_back(first + k);      // errors resurface in the narrow pass
            }
        }
    }
    if (wide.empty()) return BSW_ERANGE;               // refused for another reason: the first pass's message stands
    if (!narrow.empty()) {
        const SubsetSrc sub{ &src, narrow.data() };
        std::vector<bsw_result> r2(narrow.size());
        std::vector<uint32_t> c2(cells ? narrow.size() : 0);
        rc = run_with_reruns(ctx, cs, params, TaskSource{ &sub, fill_subset }, narrow.size(), r2.data(), cells ? c2.data() : nullptr);
        if (rc) return rc;
        for (size_t k = 0; k < narrow.size(); ++k) { out[narrow[k]] = r2[k]; if (cells) cells[narrow[k]] = c2[k]; }
    }
    return run_wide(ctx, params, src, wide.data(), wide.size(), out, cells);
}

// ---- task sources ----
void fill_flat(const void* self, size_t first, size_t count, ExtTask* out)
{
    const FlatSrc& S = *static_cast<const FlatSrc*>(self);
    const bsw_params* p = S.p;
    const BandClamp& clamp = *S.clamp;
    (void)p;
    for (size_t k = 0; k < count; ++k) {
        const size_t i = first + k;
        ExtTask& x = out[k];
        const int64_t ql = S.qoff[i + 1] - S.qoff[i], tl = S.toff[i + 1] - S.toff[i];
        x.q = S.qbuf + S.qoff[i]; x.t = S.tbuf + S.toff[i];
        x.qlen = (ql < 0 || ql > 0x7fffffff) ? -1 : (int32_t)ql;
        x.tlen = (tl < 0 || tl > 0x7fffffff) ? -1 : (int32_t)tl;
        x.h0 = S.h0[i];
        x.w = (x.qlen >= 1 && S.w[i] >= 0) ? clamp(x.qlen, S.w[i]) : -1;
    }
}
struct RecSrc { const bsw_params* p; const BandClamp* clamp; const bsw_task* tasks; };
void fill_records(const void* self, size_t first, size_t count, ExtTask* out)
{
    const RecSrc& S = *static_cast<const RecSrc*>(self);
    for (size_t k = 0; k < count; ++k) {
        const bsw_task& t = S.tasks[first + k];
        ExtTask& x = out[k];
        x.q = t.query; x.t = t.target; x.qlen = t.qlen; x.tlen = t.tlen; x.h0 = t.h0;
        x.w = (t.qlen >= 1 && t.w >= 0) ? (*S.clamp)(t.qlen, t.w) : -1;
    }
}
void fill_vector(const void* self, size_t first, size_t count, ExtTask* out)
{
    memcpy(out, static_cast<const ExtTask*>(self) + first, count * sizeof(ExtTask));
}

bool host_range_registered(bsw_ctx* ctx, const unsigned char* p, size_t bytes)
{
    std::lock_guard<std::mutex> g(ctx->err_mu);
    for (const auto& r : ctx->host_regs)
        if (p >= r.first && p + bytes <= r.first + r.second) return true;
    return false;
}

}  // namespace

// ======================================================================== C ABI

extern "C" {

const char* bsw_version(void) { return "bsw-b200 0.1 (sm_100a)"; }

int bsw_init(bsw_ctx** out, const int* device_ids, int n_devices, int streams_per_device)
{
    if (!out) return BSW_EINVAL;
    *out = nullptr;
    // Every worker slot owns streams; with the default 8 hardware queues their launches falsely serialise (measured:
    // e2e 9.2 -> 8.3 ms on 1 M tasks with 32).  Read by the driver when the context is created, so this only takes
    // effect if no CUDA context exists yet in the process; a value set by the user wins.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int ndev_avail = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev_avail);
    if (e != cudaSuccess || ndev_avail <= 0) return BSW_ECUDA;       // no CPU fallback: no device, no context
    std::unique_ptr<bsw_ctx> ctx(new bsw_ctx());
    std::vector<int> ids;
    if (!device_ids || n_devices <= 0) {
        int cur = 0;
        if (cudaGetDevice(&cur) != cudaSuccess) return BSW_ECUDA;
        ids.push_back(cur);
    } else {
        for (int k = 0; k < n_devices; ++k) {
            if (device_ids[k] < 0 || device_ids[k] >= ndev_avail) return BSW_EINVAL;
            ids.push_back(device_ids[k]);
        }
    }
    if (streams_per_device <= 0) streams_per_device = 2;
    if (streams_per_device > 8) streams_per_device = 8;
    int prev = 0;
    cudaGetDevice(&prev);
    for (int id : ids) {
        if (cudaSetDevice(id) != cudaSuccess) return BSW_ECUDA;
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, id) != cudaSuccess) return BSW_ECUDA;
        if (prop.major < 10) return BSW_ECUDA;                       // kernels are built for sm_100a only
        ctx->devs.emplace_back();
        ctx->devs.back().id = id;
        if (slot_init(ctx.get(), ctx->devs.back().aux)) { bsw_destroy(ctx.release()); cudaSetDevice(prev); return BSW_ECUDA; }
    }
    cudaSetDevice(prev);
    ctx->streams_per_device = streams_per_device;
    ctx->async.resize(16);
    *out = ctx.release();
    return BSW_OK;
}

void bsw_destroy(bsw_ctx* ctx)
{
    if (!ctx) return;
    {
        std::unique_lock<std::mutex> g(ctx->async_mu);                    // queued batches still run; then the executor stops
        ctx->async_stop = true;
    }
    ctx->async_cv.notify_all();
    for (auto& th : ctx->async_threads) th.join();
    int prev = 0;
    cudaGetDevice(&prev);
    for (auto& cs : ctx->call_free)
        for (auto& W : cs->workers) {
            cudaSetDevice(ctx->devs[(size_t)W->dev].id);
            for (Slot& s : W->slots) { if (s.stream) cudaStreamSynchronize(s.stream); slot_free(s); }
        }
    for (Device& D : ctx->devs) {
        cudaSetDevice(D.id);
        if (D.aux.stream) cudaStreamSynchronize(D.aux.stream);
        slot_free(D.aux);
    }
    for (const auto& r : ctx->host_regs) cudaHostUnregister(const_cast<unsigned char*>(r.first));
    if (ctx->trace_ref) cudaEventDestroy(ctx->trace_ref);
    cudaSetDevice(prev);
    delete ctx;
}

const char* bsw_last_error(const bsw_ctx* ctx)
{
    if (!ctx) return "null context";
    // this thread's own string (no other thread touches it); a failing call copies the text its workers left in the
    // context-wide slot into it before it returns (adopt_worker_error)
    if (tl_last_error.empty()) {
        bsw_ctx* c = const_cast<bsw_ctx*>(ctx);
        std::lock_guard<std::mutex> g(c->err_mu);
        tl_last_error = c->last_error;
    }
    return tl_last_error.c_str();
}

int bsw_num_devices(const bsw_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }

int bsw_set_option(bsw_ctx* ctx, const char* key, int64_t value)
{
    if (!ctx || !key) return BSW_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    const std::string k(key);
    if (k == "variant") { if (value != 1 && value != 2) return BSW_EINVAL; ctx->opt.variant = (int)value; }
    else if (k == "host_threads") { if (value < 0 || value > 1024) return BSW_EINVAL; ctx->opt.host_threads = (int)value; }
    else if (k == "raw_inputs") { if (value < 0 || value > 3) return BSW_EINVAL; ctx->raw_inputs = (int)value; }
    else if (k == "slots") { if (value < 1 || value > 16) return BSW_EINVAL; ctx->slots_per_worker = (int)value; }
    else if (k == "chunk_tasks") { if (value < 32) return BSW_EINVAL; ctx->chunk_tasks = (size_t)value; }
    else if (k == "force_kernel") { if (value < 0 || value > 2) return BSW_EINVAL; ctx->opt.force_kernel = (int)value; }
    else if (k == "k2_warps") { if (value != 1 && value != 4) return BSW_EINVAL; ctx->k2_warps = (int)value; }
    else if (k == "fused_l2") { ctx->fused_l2 = value != 0; }
    else if (k == "device_plan") { ctx->device_plan = value != 0; }
    else if (k == "k2_narrow") { ctx->k2_narrow = value != 0; }
    else if (k == "fpga_strict") { ctx->fpga_strict = value != 0; }
    else if (k == "wide") { if (value < 0 || value > 2) return BSW_EINVAL; ctx->wide = (int)value; }
    else if (k == "kernel_timing") { ctx->kernel_timing = value != 0; }
    else if (k == "k2_min_qlen") { if (value < 1) return BSW_EINVAL; ctx->opt.k2_min_qlen = (int)value; }
    else { set_error(ctx, "unknown option " + k); return BSW_EINVAL; }
    return BSW_OK;
}

int bsw_extend_batch(bsw_ctx* ctx, const bsw_params* params, const bsw_task* tasks, size_t n, bsw_result* out)
{
    if (!ctx) return BSW_EINVAL;
    if (n == 0) return BSW_OK;
    if (!params || !tasks || !out) { set_error(ctx, "null argument"); return BSW_EINVAL; }
    if (params->e_ins < 1 || params->e_del < 1) { set_error(ctx, "gap extension must be >= 1"); return BSW_EINVAL; }
    const BandClamp clamp(params->mat, params->end_bonus, params->o_ins, params->e_ins, params->o_del, params->e_del);
    const RecSrc S{ params, &clamp, tasks };
    return run_extensions(ctx, params, TaskSource{ &S, fill_records }, n, out, nullptr);
}

int bsw_extend_batch_flat(bsw_ctx* ctx, const bsw_params* params, const uint8_t* qbuf, const int64_t* qoff,
                          const uint8_t* tbuf, const int64_t* toff, const int32_t* h0, const int32_t* w, size_t n,
                          bsw_result* out, uint32_t* cells)
{
    if (!ctx) return BSW_EINVAL;
    if (n == 0) return BSW_OK;
    if (!params || !qbuf || !qoff || !tbuf || !toff || !h0 || !w || !out) { set_error(ctx, "null argument"); return BSW_EINVAL; }
    if (params->e_ins < 1 || params->e_del < 1) { set_error(ctx, "gap extension must be >= 1"); return BSW_EINVAL; }
    const BandClamp clamp(params->mat, params->end_bonus, params->o_ins, params->e_ins, params->o_del, params->e_del);
    const FlatSrc S{ params, &clamp, qbuf, qoff, tbuf, toff, h0, w };
    TaskSource src{ &S, fill_flat };
    // bases in registered host memory: the DMA engine reads them in place and the device packs them (raw mode)
    if (qoff[n] >= qoff[0] && toff[n] >= toff[0])
        src.raw = host_range_registered(ctx, qbuf + qoff[0], (size_t)(qoff[n] - qoff[0])) &&
                  host_range_registered(ctx, tbuf + toff[0], (size_t)(toff[n] - toff[0]));
    src.flat = &S;
    src.out_registered = host_range_registered(ctx, reinterpret_cast<const unsigned char*>(out), n * sizeof(bsw_result));
    return run_extensions(ctx, params, src, n, out, cells);
}

/* Registered host memory: page-locked in place (cudaHostRegister), so that batches whose bases live there are copied by
 * the DMA engine without the staging pass.  Ranges are remembered per context. */
int bsw_host_register(bsw_ctx* ctx, const void* ptr, size_t bytes)
{
    if (!ctx || !ptr || !bytes) return BSW_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(ctx->devs[0].id);
    const cudaError_t e = cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterPortable);
    cudaSetDevice(prev);
    if (e != cudaSuccess) { cudaGetLastError(); return cuda_fail(ctx, e, "cudaHostRegister"); }
    std::lock_guard<std::mutex> g(ctx->err_mu);
    ctx->host_regs.emplace_back(static_cast<const unsigned char*>(ptr), bytes);
    return BSW_OK;
}

int bsw_host_unregister(bsw_ctx* ctx, const void* ptr)
{
    if (!ctx || !ptr) return BSW_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    {
        std::lock_guard<std::mutex> g(ctx->err_mu);
        auto it = std::find_if(ctx->host_regs.begin(), ctx->host_regs.end(),
                               [&](const std::pair<const unsigned char*, size_t>& r) { return r.first == ptr; });
        if (it == ctx->host_regs.end()) { ctx->last_error = "bsw_host_unregister: not a registered range"; return BSW_EINVAL; }
        ctx->host_regs.erase(it);
    }
    const cudaError_t e = cudaHostUnregister(const_cast<void*>(ptr));
    if (e != cudaSuccess) { cudaGetLastError(); return cuda_fail(ctx, e, "cudaHostUnregister"); }
    return BSW_OK;
}

// ---------------- level 2: fused seed task (left + right extension, band retry, clip) ----------------
// Host-orchestrated in four device passes: left try 0, left try 1 (the few tasks whose max_off asks for the doubled
// band, sw_pe_array_sw_extend.v:1963,1824-1825,1969-1970), right try 0 with h0 = left score
// (sw_pe_array_proc_element.v:1671,1652), right try 1.  The clip decision (pe:1672-1675) runs on the host.
#define BSW_MAX_BAND_TRY 2

static int chain2aln_host(bsw_ctx* ctx, const bsw_params2* P, const bsw_seed_task* tasks, size_t n,
                          const bsw_seed_clamp* clamps, bsw_aln_record* out)
{
    if (n == 0) return BSW_OK;
    struct St { int sc0, score, truesc, qb, qe, rb, re, aw[2]; };
    std::vector<St> st(n);
    for (size_t i = 0; i < n; ++i) {                                      // pe:471-475,581-583,...,783-797
        const bsw_seed_task& s = tasks[i];
        if (s.qlen[0] < 0 || s.qlen[1] < 0) { set_error(ctx, "negative flank length"); return BSW_EINVAL; }
        st[i] = St{ s.init_score, 0, s.init_score, 0, s.qlen[1], 0, 0, { P->w, P->w } };
    }
    std::vector<ExtTask> ext;
    std::vector<size_t> who;
    std::vector<bsw_result> res, cur(n);
    std::vector<int> a_score(n), prev(n);
    for (int side = 0; side < 2; ++side) {                                // pe:1597,1622
        bsw_params pp = P->p;
        const int pen_clip = side ? P->pen_clip3 : P->pen_clip5;
        pp.end_bonus = pen_clip;                                          // BWA passes pen_clip5/3 as ksw_extend2's end_bonus
        std::vector<size_t> active;
        for (size_t i = 0; i < n; ++i)
            if (tasks[i].qlen[side] > 0) { active.push_back(i); a_score[i] = st[i].sc0; }      // pe:1670
        for (int k = 0; k < BSW_MAX_BAND_TRY && !active.empty(); ++k) {   // sx:1963,1878
            ext.clear(); who.clear();
            const int aw = P->w << k;                                     // sx:1765
            for (size_t i : active) {
                const bsw_seed_task& s = tasks[i];
                ExtTask x;
                x.q = side ? s.q_right : s.q_left; x.t = side ? s.t_right : s.t_left;
                x.qlen = s.qlen[side]; x.tlen = s.tlen[side];
                x.h0 = side ? st[i].sc0 : s.h0;                           // pe:1671,1652
                if (clamps)       // wire format: max_ins/max_del arrive precomputed from the host (pe:924-934; sx:1763-1765)
                    x.w = std::min(aw, std::min(clamps[i].max_ins[side], clamps[i].max_del[side]));
                else
                    x.w = clamp_band(pp.mat, x.qlen, aw, pp.end_bonus, pp.o_ins, pp.e_ins, pp.o_del, pp.e_del);
                ext.push_back(x); who.push_back(i);
            }
            res.resize(ext.size());
            const int rc = run_extensions(ctx, &pp, TaskSource{ ext.data(), fill_vector }, ext.size(), res.data(), nullptr);
            if (rc) return rc;
            std::vector<size_t> again;
            for (size_t e = 0; e < who.size(); ++e) {
                const size_t i = who[e];
                prev[i] = a_score[i];                                     // sx:1822,1859
                if (clamps && k > 0) {                                    // wire tasks: the FPGA carries the first try's maxima
                    const bsw_result first = cur[i];                      // (k3_rtl_carry, bsw_k3_core.cuh)
                    bsw_result& second = res[e];
                    if (!(second.score > first.score)) { second.score = first.score; second.qle = first.qle; second.tle = first.tle; }
                    if (second.gscore < first.gscore) { second.gscore = first.gscore; second.gtle = first.gtle; }
                }
                a_score[i] = res[e].score; cur[i] = res[e]; st[i].aw[side] = aw;
                if (!(a_score[i] == prev[i] || res[e].max_off < (aw >> 1) + (aw >> 2))) again.push_back(i);   // sx:1824-1825,1969-1970,1837
            }
            active.swap(again);
        }
        for (size_t i = 0; i < n; ++i) {
            const bsw_seed_task& s = tasks[i];
            if (s.qlen[side] <= 0) continue;
            St& S = st[i];
            const bsw_result& r = cur[i];
            const int sc = a_score[i];
            if (r.gscore <= 0 || r.gscore <= sc - pen_clip) {             // local: pe:1672,1674-1675,1667
                if (side == 0) { S.qb = s.qbeg - r.qle; S.rb = -r.tle; S.truesc = sc; }          // pe:591-599,659-667,767-777
                else           { S.qe = r.qle; S.re = r.tle; S.truesc += sc - S.sc0; }            // pe:615-623,683-691,1679-1680
            } else {                                                      // to-end
                if (side == 0) { S.qb = 0; S.rb = -r.gtle; S.truesc = r.gscore; }
                else           { S.qe = s.qlen[1]; S.re = r.gtle; S.truesc += r.gscore - S.sc0; }
            }
            S.score = S.sc0 = sc;                                         // pe:697-700,727-728,1594,1685
        }
    }
    for (size_t i = 0; i < n; ++i) {                                      // pe:1187-1205,1662-1665
        const St& S = st[i];
        bsw_aln_record& r = out[i];
        r.id = tasks[i].id; r.qb = S.qb; r.qe = S.qe; r.rb = S.rb; r.re = S.re;
        r.score = S.score; r.truesc = S.truesc; r.w = S.aw[0] > S.aw[1] ? S.aw[0] : S.aw[1];      // pe:1669,1684
    }
    return BSW_OK;
}

// ---- level 2, fused on the device (K3): one seed per lane, no host round trip between the two extensions ----
// Seeds whose flanks do not fit a K1 tile are returned in `leftover` for the host-orchestrated path above.
static int chain2aln_fused(bsw_ctx* ctx, const bsw_params2* P, const bsw_seed_task* tasks, size_t n,
                           const bsw_seed_clamp* clamps, bsw_aln_record* out, std::vector<size_t>* leftover)
{
    CallLease lease(ctx);
    bsw_ctx::CallState& cs = *lease.cs;
    const double call0 = now_ms();
    DevParams dp; int sym = 0, max_mat = 0; bool fast_ok = false;
    int rc = make_dev_params(ctx, &P->p, &dp, &sym, &fast_ok, &max_mat);
    if (rc) return rc;
    SchedOptions opt = ctx->opt;
    opt.fast_matrix = fast_ok; opt.force_kernel = 1;
    if (opt.host_threads <= 0) opt.host_threads = default_host_threads();
    // Validation and eligibility are checked by the workers, chunk by chunk (a serial scan of 1 M seed records cost
    // 8 ms before the first chunk was even packed): a chunk is a range of the caller's seed array, its eligible seeds
    // go through K3, the others are collected in `leftover` for the host-orchestrated path.
    std::mutex left_mu;
    auto check_seed = [&](size_t i, bool* eligible) -> int {
        const bsw_seed_task& s = tasks[i];
        const int ql = s.qlen[0], qr = s.qlen[1];
        if (ql < 0 || qr < 0 || (ql > 0 && (s.tlen[0] < 1 || s.h0 < 1 || !s.q_left || !s.t_left)) ||
            (qr > 0 && (s.tlen[1] < 1 || !s.q_right || !s.t_right)) || (ql == 0 && qr > 0 && s.init_score < 1)) {
            set_error(ctx, "seed task " + std::to_string(i) + ": empty target, h0 < 1 or null flank pointer");
            return BSW_EINVAL;
        }
        const int64_t hmax = (int64_t)std::max(s.h0, s.init_score) + (int64_t)(ql + qr) * max_mat;
        // beyond the 16-bit row state of the fused kernel: the host-orchestrated path (its extension calls split off K5 tasks)
        *eligible = !(hmax > SCORE_CAP || ql > K1_QLEN_CAP || qr > K1_QLEN_CAP || s.tlen[0] > 500000 || s.tlen[1] > 500000);
        return 0;
    };

    // the same worker-pipeline structure as run_extensions: chunks of seeds pulled from a shared counter, every worker
    // packs / plans / submits on its own two stream slots
    const size_t chunk = 8192;           // measured on 200 k seeds: 4096 -> 9.1 ms, 8192 -> 6.9 ms, 16384 -> 7.2 ms
    const size_t nchunks = (n + chunk - 1) / chunk;
    const size_t ndev = ctx->devs.size();
    size_t nworkers = std::min<size_t>((size_t)opt.host_threads, nchunks);
    if (nworkers < ndev && nchunks >= ndev) nworkers = ndev;
    if (nworkers < 1) nworkers = 1;
    for (size_t k = 0; k < nworkers; ++k) get_worker(ctx, cs, k);
    std::atomic<size_t> next(0);
    std::atomic<int> first_err(0);
    std::atomic<uint64_t> launches(0), h2d(0), d2h(0), fused_seeds(0);

    auto collect = [&](Slot& sl) -> int {
        if (!sl.busy) return 0;
        CUDA_TRY(ctx, cudaEventSynchronize(sl.ev_done));
        sl.busy = false;
        const bsw_aln_record* recs = reinterpret_cast<const bsw_aln_record*>(sl.h_out);
        const size_t nlanes = sl.plan.lane_seed.size();
        for (size_t q = 0; q < nlanes; ++q) {
            const int64_t sidx = sl.plan.lane_seed[q];
            if (sidx < 0) continue;
            out[sl.seed_idx[(size_t)sidx]] = recs[q];
        }
        return 0;
    };
    auto submit = [&](Slot& sl, size_t first, size_t range) -> int {
        int r = 0;
        sl.raw_mode = false;                     // the slot may have carried a raw-mode level-1 chunk before
        sl.dp_mode = false;                      // ... or a device-planned one (the slot accessors follow the flag)
        sl.seed_idx.clear();
        {
            std::vector<size_t> left;
            for (size_t i = first; i < first + range; ++i) {
                bool ok = false;
                if ((r = check_seed(i, &ok))) return r;
                if (ok) sl.seed_idx.push_back(i); else left.push_back(i);
            }
            if (!left.empty()) { std::lock_guard<std::mutex> g(left_mu); leftover->insert(leftover->end(), left.begin(), left.end()); }
        }
        const size_t cnt = sl.seed_idx.size();
        if (cnt == 0) return 0;
        sl.tasks.resize(2 * cnt); sl.cls.resize(2 * cnt); sl.src.resize(2 * cnt);
        for (size_t k = 0; k < cnt; ++k) {
            const bsw_seed_task& s = tasks[sl.seed_idx[k]];
            ExtTask& l = sl.tasks[2 * k]; ExtTask& rr = sl.tasks[2 * k + 1];
            l.q = s.q_left; l.t = s.t_left; l.qlen = s.qlen[0]; l.tlen = s.qlen[0] ? s.tlen[0] : 0; l.h0 = s.qlen[0] ? s.h0 : 0; l.w = s.qlen[0] ? 0 : -2;
            rr.q = s.q_right; rr.t = s.t_right; rr.qlen = s.qlen[1]; rr.tlen = s.qlen[1] ? s.tlen[1] : 0; rr.h0 = s.qlen[1] ? 1 : 0; rr.w = s.qlen[1] ? (s.qlen[0] ? std::max(s.h0, 0) + s.qlen[0] : std::max(std::max(s.init_score, s.h0), 0)) : -2;   // present flank: score-budget hint for the seed plan's sort key
        }
        const size_t src_bound = source_arena_bound(sl.tasks.data(), 2 * cnt) * 4;
        const size_t max_slots = 2 * (cnt + 2 * TILE_LANES), max_tiles = max_slots / TILE_LANES + 4;
        const size_t in_bound = src_bound + max_tiles * sizeof(TileHdr) + max_slots * (sizeof(SlotParam) + sizeof(SlotSrc)) +
                                (max_slots / 2) * sizeof(SeedParam) + 128;
        if ((r = grow_pinned(ctx, &sl.h_in, &sl.h_in_cap, in_bound))) return r;
        size_t bad = 0; std::string msg;
        r = pack_tasks(sl.tasks.data(), 2 * cnt, max_mat, opt, sl.cls.data(), sl.src.data(), reinterpret_cast<uint32_t*>(sl.h_in),
                       &sl.src_words, &bad, &msg);
        if (r) {
            set_error(ctx, "seed task " + std::to_string(sl.seed_idx[bad / 2]) + (bad & 1 ? " (right flank)" : " (left flank)") +
                               ": invalid base code or length");
            return r;
        }
        build_seed_plan(sl.tasks.data(), sl.cls.data(), sl.src.data(), cnt, opt, &sl.plan);
        Plan& PL = sl.plan;
        const size_t nslots = PL.slots.size(), nlanes = PL.lane_seed.size();
        sl.off_tiles = (sl.src_words * 4 + 15) & ~(size_t)15;
        sl.off_slots = sl.off_tiles + PL.tiles.size() * sizeof(TileHdr);
        sl.off_ssrc = sl.off_slots + nslots * sizeof(SlotParam);
        const size_t off_seeds = (sl.off_ssrc + nslots * sizeof(SlotSrc) + 15) & ~(size_t)15;
        sl.in_bytes = off_seeds + nlanes * sizeof(SeedParam);
        if (sl.in_bytes > sl.h_in_cap) { set_error(ctx, "internal: input block bound exceeded"); return BSW_ENOMEM; }
        if ((r = grow_pinned(ctx, &sl.h_out, &sl.h_out_cap, nlanes))) return r;
        if ((r = grow_device(ctx, &sl.d_in, &sl.d_in_cap, in_bound))) return r;
        if ((r = grow_device(ctx, &sl.d_arena, &sl.d_arena_cap, PL.tiled_words))) return r;
        if ((r = grow_device(ctx, &sl.d_out, &sl.d_out_cap, nlanes))) return r;
        memcpy(sl.h_in + sl.off_tiles, PL.tiles.data(), PL.tiles.size() * sizeof(TileHdr));
        memcpy(sl.h_in + sl.off_slots, PL.slots.data(), nslots * sizeof(SlotParam));
        memcpy(sl.h_in + sl.off_ssrc, PL.slot_src.data(), nslots * sizeof(SlotSrc));
        SeedParam* sp = reinterpret_cast<SeedParam*>(sl.h_in + off_seeds);
        for (size_t q = 0; q < nlanes; ++q) {
            const int64_t sidx = PL.lane_seed[q];
            SeedParam& d = sp[q];
            if (sidx < 0) { d = SeedParam{ 0, 0, -1, 0, { -1, -1 }, { -1, -1 } }; continue; }
            const size_t gi = sl.seed_idx[(size_t)sidx];
            const bsw_seed_task& s = tasks[gi];
            d.init_score = s.init_score; d.qbeg = s.qbeg; d.h0 = s.h0 > 0 ? s.h0 : 0; d.id = s.id;
            for (int side = 0; side < 2; ++side) {
                d.max_ins[side] = clamps ? clamps[gi].max_ins[side] : -1;
                d.max_del[side] = clamps ? clamps[gi].max_del[side] : -1;
            }
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(sl.d_in, sl.h_in, sl.in_bytes, cudaMemcpyHostToDevice, sl.stream));
        if ((r = enqueue_gather(ctx, sl))) return r;
        uint64_t nl = 1;
        for (const Launch& L : PL.launches) {
            LaunchArgs a{};
            a.tiles = sl.d_tiles() + L.tile0; a.slots = sl.d_slots(); a.arena = sl.d_arena;
            a.out = sl.d_out + (size_t)(L.tile0 / 2) * TILE_LANES;
            a.cells_total = nullptr; a.p = dp; a.ntiles = L.ntiles; a.qmax = L.qmax; a.nqw_max = L.nqw_max;
            a.seeds = reinterpret_cast<const SeedParam*>(sl.d_in + off_seeds) + (size_t)(L.tile0 / 2) * TILE_LANES;
            a.w = P->w; a.pen_clip5 = P->pen_clip5; a.pen_clip3 = P->pen_clip3;
            cudaError_t ce = k3_launch(a, opt.variant, L.generic, sym, sl.stream);
            if (ce != cudaSuccess) return cuda_fail(ctx, ce, "K3 launch");
            ++nl;
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(sl.h_out, sl.d_out, nlanes * sizeof(SlotResult), cudaMemcpyDeviceToHost, sl.stream));
        CUDA_TRY(ctx, cudaEventRecord(sl.ev_done, sl.stream));
        sl.busy = true; sl.first = first; sl.count = cnt;
        launches += nl; h2d += sl.in_bytes; d2h += nlanes * sizeof(SlotResult); fused_seeds += cnt;
        return 0;
    };
    auto worker_main = [&](size_t k) {
        Worker& W = *cs.workers[k];
        int r = 0;
        if (cudaSetDevice(ctx->devs[(size_t)W.dev].id) != cudaSuccess) { r = BSW_ECUDA; set_error(ctx, "cudaSetDevice failed"); }
        if (!r && W.slots.size() < (size_t)ctx->slots_per_worker) {
            const size_t have = W.slots.size();
            W.slots.resize((size_t)ctx->slots_per_worker);
            for (size_t q = have; q < W.slots.size() && !r; ++q) r = slot_init(ctx, W.slots[q]);
        }
        const size_t nslot = (size_t)ctx->slots_per_worker;
        size_t cur = 0;
        while (!r && !first_err.load(std::memory_order_relaxed)) {
            const size_t c = next.fetch_add(1);
            if (c >= nchunks) break;
            Slot& sl = W.slots[cur];
            cur = (cur + 1) % nslot;
            const double t0 = now_ms();
            if ((r = collect(sl))) break;
            const size_t first = c * chunk;
            const double t1 = now_ms();
            r = submit(sl, first, std::min(chunk, n - first));
            if (getenv("BSW_TRACE")) fprintf(stderr, "l2 w%zu chunk %zu at %.2f: collect %.2f submit %.2f ms\n", k, c, t0 - call0, t1 - t0, now_ms() - t1);
        }
        const double t2 = now_ms();
        for (Slot& sl : W.slots) {
            if (r) { if (sl.stream) cudaStreamSynchronize(sl.stream); sl.busy = false; }
            else r = collect(sl);
        }
        if (getenv("BSW_TRACE")) fprintf(stderr, "l2 w%zu drain at %.2f: %.2f ms\n", k, t2 - call0, now_ms() - t2);
        if (r) { int expect = 0; first_err.compare_exchange_strong(expect, r); }
    };
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    run_workers(nworkers, worker_main);
    cudaSetDevice(prev_dev);
    {
        std::lock_guard<std::mutex> g(ctx->err_mu);
        ctx->stats.kernel_launches += launches.load(); ctx->stats.h2d_bytes += h2d.load(); ctx->stats.d2h_bytes += d2h.load();
        ctx->stats.tasks += fused_seeds.load();
    }
    if (first_err.load()) adopt_worker_error(ctx);
    return first_err.load();
}

int bsw_chain2aln_impl(bsw_ctx* ctx, const bsw_params2* P, const bsw_seed_task* tasks, size_t n,
                       const bsw_seed_clamp* clamps, bsw_aln_record* out)
{
    if (!ctx) return BSW_EINVAL;
    if (n == 0) return BSW_OK;
    if (!P || !tasks || !out) { set_error(ctx, "null argument"); return BSW_EINVAL; }
    if (P->p.e_ins < 1 || P->p.e_del < 1 || P->w < 0) { set_error(ctx, "bad parameters"); return BSW_EINVAL; }
    if (!ctx->fused_l2) return chain2aln_host(ctx, P, tasks, n, clamps, out);
    std::vector<size_t> leftover;
    int rc = chain2aln_fused(ctx, P, tasks, n, clamps, out, &leftover);
    if (rc || leftover.empty()) return rc;
    std::sort(leftover.begin(), leftover.end());          // the workers append in completion order
    // flanks longer than a K1 tile: host-orchestrated passes (K2 does the long extensions)
    std::vector<bsw_seed_task> lt(leftover.size());
    std::vector<bsw_seed_clamp> lc(clamps ? leftover.size() : 0);
    std::vector<bsw_aln_record> lo(leftover.size());
    for (size_t k = 0; k < leftover.size(); ++k) { lt[k] = tasks[leftover[k]]; if (clamps) lc[k] = clamps[leftover[k]]; }
    rc = chain2aln_host(ctx, P, lt.data(), lt.size(), clamps ? lc.data() : nullptr, lo.data());
    if (rc) return rc;
    for (size_t k = 0; k < leftover.size(); ++k) out[leftover[k]] = lo[k];
    return BSW_OK;
}

int bsw_chain2aln_batch(bsw_ctx* ctx, const bsw_params2* P, const bsw_seed_task* tasks, size_t n, bsw_aln_record* out)
{
    return bsw_chain2aln_impl(ctx, P, tasks, n, nullptr, out);
}


// ---------------- banded global alignment with traceback (ksw_global2) ----------------
// The DP that follows seed extension in BWA-MEM (bwa_gen_cigar2 -> ksw_global2, once per reported alignment).  A plain
// driver: chunks bounded by the traceback workspace (one direction byte per cell), one stream, tiles of 32 tasks in
// input order -- the call is far from the hot path (one alignment per read against tens of extensions).
int bsw_global_batch(bsw_ctx* ctx, const bsw_params* params, const bsw_global_task* tasks, size_t n, int max_ops,
                     int32_t* score, int32_t* n_cigar, uint32_t* cigar)
{
    if (!ctx) return BSW_EINVAL;
    if (n == 0) return BSW_OK;
    if (!params || !tasks || !score || !n_cigar || !cigar || max_ops < 1) { set_error(ctx, "null argument"); return BSW_EINVAL; }
    DevParams dp; int sym = 0, max_mat = 0; bool fast_ok = false;
    int rc = make_dev_params(ctx, params, &dp, &sym, &fast_ok, &max_mat);
    if (rc) return rc;
    for (size_t i = 0; i < n; ++i) {
        const bsw_global_task& t = tasks[i];
        if (!t.query || !t.target || t.qlen < 1 || t.tlen < 1 || t.w < 0 || t.qlen > 100000 || t.tlen > 100000) {
            set_error(ctx, "global task " + std::to_string(i) + ": bad length, band or null sequence"); return BSW_EINVAL;
        }
        // bwa_gen_cigar2 widens the band to at least the length difference before it calls ksw_global2: a narrower band
        // cannot reach the last cell and the traceback would leave the stored band
        if (t.w < std::abs(t.tlen - t.qlen)) { set_error(ctx, "global task " + std::to_string(i) + ": band narrower than |tlen - qlen|"); return BSW_EINVAL; }
        for (int k = 0; k < t.qlen; ++k) if (t.query[k] > 4) { set_error(ctx, "global task " + std::to_string(i) + ": invalid (base code > 4)"); return BSW_EINVAL; }
        for (int k = 0; k < t.tlen; ++k) if (t.target[k] > 4) { set_error(ctx, "global task " + std::to_string(i) + ": invalid (base code > 4)"); return BSW_EINVAL; }
    }
    int prev = 0;
    cudaGetDevice(&prev);
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[0].id));
    cudaStream_t st = nullptr;
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const size_t z_budget = (size_t)1 << 30;                         // bytes of direction matrix per chunk
    std::vector<GlobalTask> gt;
    std::vector<uint8_t> qb, tb;
    std::vector<uint32_t> eh_off;
    std::vector<uint64_t> z_off;
    size_t first = 0;
    auto fail = [&](int code) { cudaStreamDestroy(st); cudaSetDevice(prev); return code; };
    while (first < n) {
        gt.clear(); qb.clear(); tb.clear(); eh_off.clear(); z_off.clear();
        uint64_t z_rows = 0; uint64_t eh_rows = 0;
        size_t count = 0;
        while (first + count < n && count < 65536) {
            // one tile of up to 32 tasks
            const size_t tn = std::min<size_t>(TILE_LANES, n - first - count);
            uint64_t zmax = 0; int qmax = 0;
            for (size_t k = 0; k < tn; ++k) {
                const bsw_global_task& t = tasks[first + count + k];
                const uint64_t ncol = (uint64_t)std::min(t.qlen, 2 * t.w + 1);
                zmax = std::max<uint64_t>(zmax, ncol * (uint64_t)t.tlen);
                qmax = std::max(qmax, t.qlen);
            }
            if (count > 0 && (z_rows + zmax) * TILE_LANES > z_budget) break;
            if (zmax * TILE_LANES > ((size_t)8 << 30)) { set_error(ctx, "global task " + std::to_string(first + count) + ": traceback matrix too large"); return fail(BSW_ERANGE); }
            eh_off.push_back((uint32_t)eh_rows); z_off.push_back(z_rows);
            eh_rows += 2 * (uint64_t)(qmax + 2); z_rows += zmax;
            for (size_t k = 0; k < tn; ++k) {
                const bsw_global_task& t = tasks[first + count + k];
                gt.push_back(GlobalTask{ (uint32_t)qb.size(), (uint32_t)tb.size(), t.qlen, t.tlen, t.w });
                qb.insert(qb.end(), t.query, t.query + t.qlen); tb.insert(tb.end(), t.target, t.target + t.tlen);
            }
            count += tn;
        }
        GlobalArgs a{};
        void *d_tasks = nullptr, *d_q = nullptr, *d_t = nullptr, *d_eh = nullptr, *d_z = nullptr, *d_eo = nullptr, *d_zo = nullptr, *d_sc = nullptr, *d_nc = nullptr, *d_cg = nullptr;
        auto release = [&]() { for (void* p : { d_tasks, d_q, d_t, d_eh, d_z, d_eo, d_zo, d_sc, d_nc, d_cg }) if (p) cudaFree(p); };
#define G_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { release(); int c__ = cuda_fail(ctx, e__, #call); return fail(c__); } } while (0)
        G_TRY(cudaMalloc(&d_tasks, gt.size() * sizeof(GlobalTask)));
        G_TRY(cudaMalloc(&d_q, qb.size() + 16)); G_TRY(cudaMalloc(&d_t, tb.size() + 16));
        G_TRY(cudaMalloc(&d_eh, (size_t)eh_rows * TILE_LANES * sizeof(int32_t)));
        G_TRY(cudaMalloc(&d_z, (size_t)z_rows * TILE_LANES + 16));
        G_TRY(cudaMalloc(&d_eo, eh_off.size() * sizeof(uint32_t))); G_TRY(cudaMalloc(&d_zo, z_off.size() * sizeof(uint64_t)));
        G_TRY(cudaMalloc(&d_sc, count * sizeof(int32_t))); G_TRY(cudaMalloc(&d_nc, count * sizeof(int32_t)));
        G_TRY(cudaMalloc(&d_cg, count * (size_t)max_ops * sizeof(uint32_t)));
        G_TRY(cudaMemcpyAsync(d_tasks, gt.data(), gt.size() * sizeof(GlobalTask), cudaMemcpyHostToDevice, st));
        G_TRY(cudaMemcpyAsync(d_q, qb.data(), qb.size(), cudaMemcpyHostToDevice, st));
        G_TRY(cudaMemcpyAsync(d_t, tb.data(), tb.size(), cudaMemcpyHostToDevice, st));
        G_TRY(cudaMemcpyAsync(d_eo, eh_off.data(), eh_off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        G_TRY(cudaMemcpyAsync(d_zo, z_off.data(), z_off.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
        a.tasks = static_cast<const GlobalTask*>(d_tasks); a.qbuf = static_cast<const uint8_t*>(d_q); a.tbuf = static_cast<const uint8_t*>(d_t);
        a.eh = static_cast<int32_t*>(d_eh); a.z = static_cast<uint8_t*>(d_z); a.eh_off = static_cast<const uint32_t*>(d_eo); a.z_off = static_cast<const uint64_t*>(d_zo);
        a.score = static_cast<int32_t*>(d_sc); a.n_cigar = static_cast<int32_t*>(d_nc); a.cigar = static_cast<uint32_t*>(d_cg);
        a.ntasks = (uint32_t)count; a.max_ops = max_ops; a.p = dp;
        G_TRY(k4_launch(a, st));
        G_TRY(cudaMemcpyAsync(score + first, d_sc, count * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        G_TRY(cudaMemcpyAsync(n_cigar + first, d_nc, count * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        G_TRY(cudaMemcpyAsync(cigar + first * (size_t)max_ops, d_cg, count * (size_t)max_ops * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        G_TRY(cudaStreamSynchronize(st));
#undef G_TRY
        release();
        {
            std::lock_guard<std::mutex> g(ctx->err_mu);
            ctx->stats.kernel_launches += 1; ctx->stats.tasks += count;
        }
        first += count;
    }
    cudaStreamDestroy(st);
    cudaSetDevice(prev);
    for (size_t i = 0; i < n; ++i)
        if (n_cigar[i] < 0) { set_error(ctx, "global task " + std::to_string(i) + ": the alignment needs more than max_ops CIGAR operations"); return BSW_ERANGE; }
    return BSW_OK;
}

void bsw_set_error_text(bsw_ctx* ctx, const char* text) { set_error(ctx, text ? text : ""); }
int bsw_option_value(bsw_ctx* ctx, const char* key)
{
    if (!ctx || !key) return -1;
    if (!strcmp(key, "fpga_strict")) return ctx->fpga_strict ? 1 : 0;
    return -1;
}

// ---------------- async trio ----------------
// "Write REQ_PEARRAY, poll the DSM busy bit" (batch_manager.v:347,851-854) with up to 16 tickets open.  Submitted
// batches are queued to a small persistent executor (started by the first submit; BSW_ASYNC_THREADS, default 4 = the
// FPGA's four PE arrays), each runs as an ordinary batch call on its own CallState, so up to four overlap on the GPU.
// Ticket life: submit -> (poll)* -> wait.  bsw_poll never releases the ticket (0 = running, 1 = done, < 0 = failed);
// bsw_wait blocks until the batch is done, returns its status, and releases the ticket -- always finish with it.
void bsw_ctx::async_main()
{
    for (;;) {
        std::function<int()> job; size_t slot = 0;
        {
            std::unique_lock<std::mutex> g(async_mu);
            async_cv.wait(g, [&] { return async_stop || !async_queue.empty(); });
            if (async_queue.empty()) return;                              // stop requested and nothing left to run
            slot = async_queue.front().first; job = std::move(async_queue.front().second);
            async_queue.pop_front();
        }
        const int rc = job();
        std::string text = rc ? tl_last_error : std::string();
        {
            std::lock_guard<std::mutex> g(async_mu);
            async[slot].status = rc; async[slot].error.swap(text); async[slot].done = true;
        }
        async_done_cv.notify_all();
    }
}

int bsw_submit(bsw_ctx* ctx, const bsw_params* params, const bsw_task* tasks, size_t n, bsw_result* out, bsw_ticket* ticket)
{
    if (!ctx || !ticket) return BSW_EINVAL;
    std::lock_guard<std::mutex> g(ctx->async_mu);
    if (ctx->async_threads.empty()) {
        const char* e = getenv("BSW_ASYNC_THREADS");
        const int nt = e ? std::max(1, std::min(16, atoi(e))) : 4;
        for (int k = 0; k < nt; ++k) ctx->async_threads.emplace_back([ctx]() { ctx->async_main(); });
    }
    for (size_t k = 0; k < ctx->async.size(); ++k) {
        auto& a = ctx->async[k];
        if (a.used) continue;
        a.used = true; a.done = false; a.seq = ++ctx->async_seq; a.status = 0; a.error.clear();
        const bsw_params pcopy = params ? *params : bsw_params{};
        const bool have_params = params != nullptr;
        ctx->async_queue.emplace_back(k, [=]() { return bsw_extend_batch(ctx, have_params ? &pcopy : nullptr, tasks, n, out); });
        ticket->slot = (int32_t)k; ticket->seq = a.seq;
        ctx->async_cv.notify_one();
        return BSW_OK;
    }
    set_error(ctx, "no free async slot");
    return BSW_EBUSY;
}

int bsw_poll(bsw_ctx* ctx, const bsw_ticket* t)
{
    if (!ctx || !t || t->slot < 0 || (size_t)t->slot >= ctx->async.size()) return BSW_EINVAL;
    std::lock_guard<std::mutex> g(ctx->async_mu);
    const auto& a = ctx->async[(size_t)t->slot];
    if (!a.used || a.seq != t->seq) return BSW_EINVAL;
    if (!a.done) return 0;
    return a.status == BSW_OK ? 1 : a.status;
}

int bsw_wait(bsw_ctx* ctx, const bsw_ticket* t)
{
    if (!ctx || !t || t->slot < 0 || (size_t)t->slot >= ctx->async.size()) return BSW_EINVAL;
    std::unique_lock<std::mutex> g(ctx->async_mu);
    auto& a = ctx->async[(size_t)t->slot];
    if (!a.used || a.seq != t->seq) return BSW_EINVAL;
    const uint32_t seq = a.seq;
    ctx->async_done_cv.wait(g, [&] { return a.done || !a.used || a.seq != seq; });
    if (!a.used || a.seq != seq) return BSW_EINVAL;                       // another thread finished the same ticket first
    const int rc = a.status;
    if (rc) tl_last_error = a.error;
    a.used = false;
    return rc;
}

// ---------------- device-resident batches ----------------
int bsw_resident_create(bsw_ctx* ctx, const bsw_params* params, const uint8_t* qbuf, const int64_t* qoff,
                        const uint8_t* tbuf, const int64_t* toff, const int32_t* h0, const int32_t* w, size_t n,
                        bsw_resident** out)
{
    if (!ctx || !out) return BSW_EINVAL;
    *out = nullptr;
    if (!params || !qbuf || !qoff || !tbuf || !toff || !h0 || !w || n == 0) { set_error(ctx, "null argument"); return BSW_EINVAL; }
    std::lock_guard<std::mutex> lock(ctx->mu);
    std::unique_ptr<bsw_resident> R(new bsw_resident());
    int max_mat = 0; bool fast_ok = false;
    int rc = make_dev_params(ctx, params, &R->dp, &R->sym, &fast_ok, &max_mat);
    if (rc) return rc;
    SchedOptions opt = ctx->opt;
    opt.fast_matrix = fast_ok;
    if (opt.host_threads <= 0) opt.host_threads = default_host_threads();
    std::vector<ExtTask> v(n);
    for (size_t i = 0; i < n; ++i) {
        ExtTask& x = v[i];
        x.q = qbuf + qoff[i]; x.t = tbuf + toff[i];
        x.qlen = (int32_t)(qoff[i + 1] - qoff[i]); x.tlen = (int32_t)(toff[i + 1] - toff[i]); x.h0 = h0[i];
        x.w = (x.qlen >= 1 && w[i] >= 0) ? clamp_band(params->mat, x.qlen, w[i], params->end_bonus, params->o_ins,
                                                      params->e_ins, params->o_del, params->e_del) : -1;
    }
    R->dev = ctx->devs[0].id; R->n = n; R->variant = opt.variant;
    int prev = 0;
    cudaGetDevice(&prev);
    CUDA_TRY(ctx, cudaSetDevice(R->dev));
    rc = slot_init(ctx, R->slot);
    if (!rc) {
        // submit once: this packs, uploads and runs the batch; later runs reuse the device-resident plan
        LocalStats st;
        R->slot.tasks.swap(v);
        rc = slot_submit(ctx, R->slot, 0, n, max_mat, R->dp, R->sym, opt, false, &st);
        std::vector<ExtTask>().swap(R->slot.tasks);
        if (!rc) { cudaError_t e = cudaStreamSynchronize(R->slot.stream); if (e != cudaSuccess) rc = cuda_fail(ctx, e, "resident upload"); }
        R->slot.busy = false;
    }
    cudaSetDevice(prev);
    if (rc) { slot_free(R->slot); return rc; }
    *out = R.release();
    return BSW_OK;
}

int bsw_resident_run(bsw_ctx* ctx, bsw_resident* R, double* kernel_ms, uint64_t* cells, uint64_t* launches)
{
    if (!ctx || !R) return BSW_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    int prev = 0;
    cudaGetDevice(&prev);
    CUDA_TRY(ctx, cudaSetDevice(R->dev));
    Slot& s = R->slot;
    CUDA_TRY(ctx, cudaMemsetAsync(s.d_cells, 0, sizeof(unsigned long long), s.stream));
    CUDA_TRY(ctx, cudaEventRecord(s.ev_k0, s.stream));
    size_t nl = 0;
    { int rc = enqueue_gather(ctx, s);                  // the device half of the scheduler is part of every pass
      if (!rc) rc = enqueue_launches(ctx, s, R->dp, R->sym, R->variant, true, &nl, nullptr);
      if (rc) { cudaSetDevice(prev); return rc; }
      if (s.plan.n_k1_tiles) ++nl; }
    CUDA_TRY(ctx, cudaEventRecord(s.ev_k1, s.stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(s.h_cells, s.d_cells, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(s.stream));
    float ms = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1));
    R->last_ms = ms;
    if (kernel_ms) *kernel_ms = ms;
    if (cells) *cells = *s.h_cells;
    if (launches) *launches = nl;
    {
        std::lock_guard<std::mutex> g(ctx->err_mu);
        ctx->stats.tasks += R->n; ctx->stats.cells_band += *s.h_cells; ctx->stats.kernel_ms += ms; ctx->stats.kernel_launches += nl;
    }
    cudaSetDevice(prev);
    return BSW_OK;
}

int bsw_resident_fetch(bsw_ctx* ctx, bsw_resident* R, bsw_result* out, uint32_t* cells)
{
    if (!ctx || !R || !out) return BSW_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    int prev = 0;
    cudaGetDevice(&prev);
    CUDA_TRY(ctx, cudaSetDevice(R->dev));
    Slot& s = R->slot;
    const Plan& P = s.plan;
    CUDA_TRY(ctx, cudaMemcpyAsync(s.h_out, s.d_out, P.slots.size() * sizeof(SlotResult), cudaMemcpyDeviceToHost, s.stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(s.stream));
    for (size_t k = 0; k < P.slots.size(); ++k) {
        const int64_t t = P.slot_task[k];
        if (t < 0) continue;
        const SlotResult& r = s.h_out[k];
        bsw_result& o = out[(size_t)t];
        o.score = r.score; o.qle = r.qle; o.tle = r.tle; o.gtle = r.gtle; o.gscore = r.gscore; o.max_off = r.max_off;
        if (cells) cells[(size_t)t] = (uint32_t)r.cells;
    }
    cudaSetDevice(prev);
    return BSW_OK;
}

void bsw_resident_free(bsw_ctx* ctx, bsw_resident* R)
{
    if (!R) return;
    (void)ctx;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(R->dev);
    if (R->slot.stream) cudaStreamSynchronize(R->slot.stream);
    slot_free(R->slot);
    cudaSetDevice(prev);
    delete R;
}

int bsw_get_stats(bsw_ctx* ctx, bsw_stats* out)
{
    if (!ctx || !out) return BSW_EINVAL;
    std::lock_guard<std::mutex> g(ctx->err_mu);
    *out = ctx->stats;
    return BSW_OK;
}

int bsw_reset_stats(bsw_ctx* ctx)
{
    if (!ctx) return BSW_EINVAL;
    std::lock_guard<std::mutex> g(ctx->err_mu);
    ctx->stats = bsw_stats{};
    return BSW_OK;
}

int bsw_measure_int_peak(bsw_ctx* ctx, int device_index, bsw_int_peak* out)
{
    if (!ctx || !out || device_index < 0 || (size_t)device_index >= ctx->devs.size()) return BSW_EINVAL;
    std::lock_guard<std::mutex> lock(ctx->mu);
    int prev = 0;
    cudaGetDevice(&prev);
    CUDA_TRY(ctx, cudaSetDevice(ctx->devs[(size_t)device_index].id));
    double ops[5] = { 0, 0, 0, 0, 0 }; double mhz = 0; int sms = 0;
    cudaError_t e = int_peak_run(ops, &mhz, &sms, ctx->devs[(size_t)device_index].aux.stream);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "int peak micro-benchmark");
    out->iadd_tops = ops[0] * 1e-12; out->vimnmx_tops = ops[1] * 1e-12; out->dpx_tops = ops[2] * 1e-12; out->mix_tops = ops[3] * 1e-12; out->dual_tops = ops[4] * 1e-12;
    out->sm_clock_mhz = mhz; out->sm_count = sms;
    return BSW_OK;
}

}  // extern "C"
