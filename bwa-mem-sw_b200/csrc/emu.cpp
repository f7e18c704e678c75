// emu.cpp -- TEST INFRASTRUCTURE ONLY (libbsw_emu.so).  NOT linked into libbsw.so, never used by the product path.
//
// Runs the host half of the level-1 call exactly as libbsw.so does (validate -> plan -> pack into the tiled arena)
// and then executes the K1 lane function (bsw_k1_core.cuh, the same source the device kernel instantiates) on the
// CPU, lane by lane, over emulated shared-memory arrays.  This lets `pytest -m "not gpu"` check the scheduler, the
// packer and the K1 control flow (lazy narrowing, partial chunks, DPX identities) against the oracle without a GPU.
// The CUDA kernels themselves are only ever checked on a GPU (`pytest -m gpu`).
#include <atomic>
#include <chrono>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bsw.h"
#include "bsw_device.cuh"
#include "bsw_k1_core.cuh"
#include "bsw_k3_core.cuh"
#include "bsw_sched.h"

using namespace bsw;

namespace {

template <int VARIANT, int GENERIC, int SYM>
void run_tile(const DevParams& dp, const TileHdr& hd, const SlotParam* slots, const uint32_t* arena, SlotResult* out,
              int qmax, int nqw_max)
{
    const int nqw = (int)(hd.nqw_ntw & 0x7fffu);
    std::vector<uint32_t> qs((size_t)(nqw_max + K1_QS_EXTRA) * K1_S, 0xdeadbeefu);
    std::vector<uint32_t> eh((size_t)(qmax + 1 + K1_EH_SLACK) * K1_S, 0xdeadbeefu);
    memcpy(qs.data(), arena + (size_t)hd.qoff16 * 4, (size_t)nqw * K1_S * 4);       // the TMA bulk copy
    for (int lane = 0; lane < K1_S; ++lane) {
        const SlotParam& sp = slots[hd.slot0 + lane];
        if (sp.qlen <= 0) continue;
        const uint32_t* tg = arena + (size_t)hd.toff16 * 4 + lane;
        k1_task<VARIANT, GENERIC, SYM>(dp, sp.qlen, sp.tlen, sp.h0, sp.w, nqw, eh.data() + lane, qs.data() + lane, tg,
                                       out[hd.slot0 + lane]);
    }
}

// Host restatement of the k0 gather kernel (bsw_k0.cu): lane l of a K1 tile copies its task's packed words from the
// task-major source arena into the tile-interleaved block.
void gather_host(const Plan& P, const uint32_t* src, uint32_t* dst)
{
    for (uint32_t t = 0; t < P.n_k1_tiles; ++t) {
        const TileHdr& hd = P.tiles[t];
        const int nqw = (int)(hd.nqw_ntw & 0x7fffu), ntw = (int)(hd.nqw_ntw >> 16);
        for (int lane = 0; lane < TILE_LANES; ++lane) {
            const SlotParam& sp = P.slots[hd.slot0 + lane];
            const SlotSrc& ss = P.slot_src[hd.slot0 + lane];
            const int own_q = sp.qlen > 0 ? ((sp.qlen + 31) >> 5) * 4 : 0, own_t = sp.tlen > 0 ? ((sp.tlen + 31) >> 5) * 4 : 0;
            for (int k = 0; k < nqw; ++k) dst[(size_t)hd.qoff16 * 4 + (size_t)k * TILE_LANES + lane] = k < own_q ? src[(size_t)ss.qoff16 * 4 + k] : 0u;
            for (int k = 0; k < ntw; ++k) dst[(size_t)hd.toff16 * 4 + (size_t)k * TILE_LANES + lane] = k < own_t ? src[(size_t)ss.toff16 * 4 + k] : 0u;
        }
    }
}

}  // namespace

extern "C" {

// info[0] = launches, info[1] = tiles, info[2] = arena words, info[3] = padded lanes
int bsw_emu_force_kernel = 1;   // 1: everything on K1; 0: auto (tasks that need K2 make the call fail)
int bsw_emu_k2_min_qlen = 384;

int bsw_emu_extend_batch_flat(const bsw_params* params, int variant, const uint8_t* qbuf, const int64_t* qoff,
                              const uint8_t* tbuf, const int64_t* toff, const int32_t* h0, const int32_t* w, size_t n,
                              bsw_result* out, uint32_t* cells, int64_t* info)
{
    if (!params || params->e_ins < 1 || params->e_del < 1 || params->o_ins < 0 || params->o_del < 0) return BSW_EINVAL;
    DevParams dp;
    memset(&dp, 0, sizeof(dp));
    dp.o_del = params->o_del; dp.e_del = params->e_del; dp.o_ins = params->o_ins; dp.e_ins = params->e_ins; dp.zdrop = params->zdrop;
    int mx = 0;
    for (int k = 0; k < 25; ++k) { dp.mat[k] = params->mat[k]; mx = mx > params->mat[k] ? mx : params->mat[k]; }
    bool fast = true;
    const int a = params->mat[0], b = -params->mat[1];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (params->mat[5 * i + j] != (i == j ? a : -b)) fast = false;
    dp.match = a; dp.mismatch = b;
    for (int t = 0; t < 5; ++t) {
        uint32_t lo = 0;
        for (int q = 0; q < 4; ++q) lo |= (uint32_t)(uint8_t)params->mat[5 * t + q] << (8 * q);
        dp.row_lo[t] = lo; dp.row_hi[t] = (uint32_t)(uint8_t)params->mat[5 * t + 4];
    }
    const int sym = (params->o_del == params->o_ins && params->e_del == params->e_ins) ? 1 : 0;

    SchedOptions opt;
    opt.variant = variant; opt.force_kernel = bsw_emu_force_kernel; opt.fast_matrix = fast; opt.host_threads = 4;
    opt.k2_min_qlen = bsw_emu_k2_min_qlen;
    std::vector<ExtTask> v(n);
    for (size_t i = 0; i < n; ++i) {
        ExtTask& x = v[i];
        x.q = qbuf + qoff[i]; x.t = tbuf + toff[i];
        x.qlen = (int32_t)(qoff[i + 1] - qoff[i]); x.tlen = (int32_t)(toff[i + 1] - toff[i]); x.h0 = h0[i];
        x.w = (x.qlen >= 1 && w[i] >= 0) ? clamp_band(params->mat, x.qlen, w[i], params->end_bonus, params->o_ins,
                                                      params->e_ins, params->o_del, params->e_del) : -1;
    }
    std::vector<uint8_t> cls(n);
    std::vector<SlotSrc> ssrc(n);
    std::vector<uint32_t> src(source_arena_bound(v.data(), n), 0xdeadbeefu);
    size_t bad = 0, used = 0; std::string msg;
    int rc = pack_tasks(v.data(), n, mx, opt, cls.data(), ssrc.data(), src.data(), &used, &bad, &msg);
    if (rc) return rc;
    if (used > src.size()) return BSW_ENOMEM;
    Plan P;
    build_plan(v.data(), cls.data(), ssrc.data(), n, opt, &P);
    std::vector<uint32_t> arena(P.tiled_words + 64, 0xdeadbeefu);
    gather_host(P, src.data(), arena.data());
    std::vector<SlotResult> res(P.slots.size());
    for (const Launch& L : P.launches) {
        if (L.kind != 1) return BSW_ERANGE;
        for (uint32_t t = L.tile0; t < L.tile0 + L.ntiles; ++t) {
            const TileHdr& hd = P.tiles[t];
#define EMU_CASE(V, G, S) if (variant == V && L.generic == G && sym == S) run_tile<V, G, S>(dp, hd, P.slots.data(), arena.data(), res.data(), L.qmax, L.nqw_max);
            EMU_CASE(1, 0, 1) EMU_CASE(1, 0, 0) EMU_CASE(1, 1, 1) EMU_CASE(1, 1, 0)
            EMU_CASE(2, 0, 1) EMU_CASE(2, 0, 0) EMU_CASE(2, 1, 1) EMU_CASE(2, 1, 0)
#undef EMU_CASE
        }
    }
    size_t pad = 0;
    for (size_t k = 0; k < P.slots.size(); ++k) {
        const int64_t t = P.slot_task[k];
        if (t < 0) { ++pad; continue; }
        const SlotResult& r = res[k];
        bsw_result& o = out[t];
        o.score = r.score; o.qle = r.qle; o.tle = r.tle; o.gtle = r.gtle; o.gscore = r.gscore; o.max_off = r.max_off;
        if (cells) cells[t] = (uint32_t)r.cells;
    }
    if (info) { info[0] = (int64_t)P.launches.size(); info[1] = (int64_t)P.tiles.size(); info[2] = (int64_t)P.tiled_words; info[3] = (int64_t)pad; }
    return BSW_OK;
}

}  // extern "C"

// Host-path profiler (tests/dev only): the per-chunk host pipeline of bsw_host.cpp's workers without the CUDA calls.
// ms[0..2] = fill, validate+pack, plan (cpu-ms summed over chunks), ms[4] = wall; `threads` chunks in flight.
extern "C" int bsw_emu_host_phases(const bsw_params* params, const uint8_t* qbuf, const int64_t* qoff, const uint8_t* tbuf,
                                   const int64_t* toff, const int32_t* h0, const int32_t* w, size_t n, int threads, size_t chunk,
                                   double* ms)
{
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    SchedOptions opt;
    opt.host_threads = 1; opt.fast_matrix = true;
    const size_t nchunks = (n + chunk - 1) / chunk;
    struct Scratch { std::vector<ExtTask> v; std::vector<uint8_t> cls; std::vector<SlotSrc> ss; Plan P; std::vector<uint32_t> arena; double t[4] = {0, 0, 0, 0}; };
    static std::vector<Scratch> sc;
    if (sc.size() < (size_t)threads) sc.resize((size_t)threads);
    std::atomic<size_t> next(0);
    std::atomic<int> slot(0);
    const BandClamp clamp(params->mat, params->end_bonus, params->o_ins, params->e_ins, params->o_del, params->e_del);
    const double w0 = now();
    pfor((size_t)threads, 1, threads, [&](size_t, size_t) {
        Scratch& S = sc[(size_t)slot.fetch_add(1)];
        for (int k = 0; k < 4; ++k) S.t[k] = 0;
        for (;;) {
            const size_t c = next.fetch_add(1);
            if (c >= nchunks) break;
            const size_t first = c * chunk, cnt = std::min(chunk, n - first);
            double t0 = now();
            S.v.resize(cnt); S.cls.resize(cnt); S.ss.resize(cnt);
            for (size_t k = 0; k < cnt; ++k) {
                const size_t i = first + k;
                ExtTask& x = S.v[k];
                x.q = qbuf + qoff[i]; x.t = tbuf + toff[i];
                x.qlen = (int32_t)(qoff[i + 1] - qoff[i]); x.tlen = (int32_t)(toff[i + 1] - toff[i]); x.h0 = h0[i];
                x.w = clamp(x.qlen, w[i]);
            }
            double t1 = now();
            const size_t bound = source_arena_bound(S.v.data(), cnt);
            if (S.arena.size() < bound) S.arena.resize(bound + bound / 4);
            size_t bad = 0, used = 0; std::string msg;
            pack_tasks(S.v.data(), cnt, 1, opt, S.cls.data(), S.ss.data(), S.arena.data(), &used, &bad, &msg);
            double t2 = now();
            build_plan(S.v.data(), S.cls.data(), S.ss.data(), cnt, opt, &S.P);
            double t3 = now();
            S.t[0] += t1 - t0; S.t[1] += t2 - t1; S.t[2] += t3 - t2;
        }
    });
    ms[4] = now() - w0;
    for (int k = 0; k < 4; ++k) { ms[k] = 0; for (int t = 0; t < threads; ++t) ms[k] += sc[(size_t)t].t[k]; }
    return 0;
}

// Level 2 on the CPU: the host half of bsw_chain2aln_batch's fused path (pack, seed plan, gather) + the K3 lane function.
namespace {
template <int VARIANT, int GENERIC, int SYM>
void run_seed_pair(const DevParams& dp, int w, int pc5, int pc3, const TileHdr& hl, const TileHdr& hr, const SlotParam* slots,
                   const SeedParam* seeds, const uint32_t* arena, SeedRecord* out, int qmax, int nqw_max)
{
    const int nql = (int)(hl.nqw_ntw & 0x7fffu), nqr = (int)(hr.nqw_ntw & 0x7fffu);
    const size_t qwords = (size_t)(nqw_max + K1_QS_EXTRA) * K1_S;
    std::vector<uint32_t> qsl(qwords, 0xdeadbeefu), qsr(qwords, 0xdeadbeefu);
    std::vector<uint32_t> eh((size_t)(qmax + 1 + K1_EH_SLACK) * K1_S, 0xdeadbeefu);
    memcpy(qsl.data(), arena + (size_t)hl.qoff16 * 4, (size_t)nql * K1_S * 4);
    memcpy(qsr.data(), arena + (size_t)hr.qoff16 * 4, (size_t)nqr * K1_S * 4);
    for (int lane = 0; lane < K1_S; ++lane) {
        if (seeds[lane].h0 < 0) continue;
        uint32_t cells = 0;
        k3_seed<VARIANT, GENERIC, SYM>(dp, w, pc5, pc3, slots[hl.slot0 + lane], slots[hr.slot0 + lane], seeds[lane], nql, nqr,
                                       eh.data() + lane, qsl.data() + lane, qsr.data() + lane,
                                       arena + (size_t)hl.toff16 * 4 + lane, arena + (size_t)hr.toff16 * 4 + lane, out[lane], cells);
    }
}
}  // namespace

// gaps4 (may be null): per seed {max_ins_left, max_del_left, max_ins_right, max_del_right} as a TBB carries them --
// the wire mode of bsw_fpga_batch (host-supplied band clamp, the FPGA's carried second band try).
extern "C" int bsw_emu_chain2aln_wire(const bsw_params2* P2, int variant, const bsw_seed_task* tasks, size_t n,
                                      const int32_t* gaps4, bsw_aln_record* out);

extern "C" int bsw_emu_chain2aln(const bsw_params2* P2, int variant, const bsw_seed_task* tasks, size_t n, bsw_aln_record* out)
{
    return bsw_emu_chain2aln_wire(P2, variant, tasks, n, nullptr, out);
}

extern "C" int bsw_emu_chain2aln_wire(const bsw_params2* P2, int variant, const bsw_seed_task* tasks, size_t n,
                                      const int32_t* gaps4, bsw_aln_record* out)
{
    const bsw_params* params = &P2->p;
    if (params->e_ins < 1 || params->e_del < 1 || params->o_ins < 0 || params->o_del < 0) return BSW_EINVAL;
    DevParams dp;
    memset(&dp, 0, sizeof(dp));
    dp.o_del = params->o_del; dp.e_del = params->e_del; dp.o_ins = params->o_ins; dp.e_ins = params->e_ins; dp.zdrop = params->zdrop;
    int mx = 0;
    for (int k = 0; k < 25; ++k) { dp.mat[k] = params->mat[k]; mx = mx > params->mat[k] ? mx : params->mat[k]; }
    dp.max_mat = mx;
    bool fast = true;
    const int a = params->mat[0], b = -params->mat[1];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (params->mat[5 * i + j] != (i == j ? a : -b)) fast = false;
    dp.match = a; dp.mismatch = b;
    for (int t = 0; t < 5; ++t) {
        uint32_t lo = 0;
        for (int q = 0; q < 4; ++q) lo |= (uint32_t)(uint8_t)params->mat[5 * t + q] << (8 * q);
        dp.row_lo[t] = lo; dp.row_hi[t] = (uint32_t)(uint8_t)params->mat[5 * t + 4];
    }
    const int sym = (params->o_del == params->o_ins && params->e_del == params->e_ins) ? 1 : 0;
    SchedOptions opt;
    opt.variant = variant; opt.force_kernel = 1; opt.fast_matrix = fast; opt.host_threads = 1;
    std::vector<ExtTask> v(2 * n);
    for (size_t k = 0; k < n; ++k) {
        const bsw_seed_task& s = tasks[k];
        ExtTask& l = v[2 * k]; ExtTask& r = v[2 * k + 1];
        l.q = s.q_left; l.t = s.t_left; l.qlen = s.qlen[0]; l.tlen = s.qlen[0] ? s.tlen[0] : 0; l.h0 = s.qlen[0] ? s.h0 : 0; l.w = s.qlen[0] ? 0 : -2;
        r.q = s.q_right; r.t = s.t_right; r.qlen = s.qlen[1]; r.tlen = s.qlen[1] ? s.tlen[1] : 0; r.h0 = s.qlen[1] ? 1 : 0; r.w = s.qlen[1] ? (s.qlen[0] ? std::max(s.h0, 0) + s.qlen[0] : std::max(std::max(s.init_score, s.h0), 0)) : -2;   // present flank: score-budget hint for the seed plan's sort key
    }
    std::vector<uint8_t> cls(2 * n);
    std::vector<SlotSrc> ssrc(2 * n);
    std::vector<uint32_t> src(source_arena_bound(v.data(), 2 * n), 0xdeadbeefu);
    size_t bad = 0, used = 0; std::string msg;
    int rc = pack_tasks(v.data(), 2 * n, mx, opt, cls.data(), ssrc.data(), src.data(), &used, &bad, &msg);
    if (rc) return rc;
    Plan P;
    build_seed_plan(v.data(), cls.data(), ssrc.data(), n, opt, &P);
    std::vector<uint32_t> arena(P.tiled_words + 64, 0xdeadbeefu);
    gather_host(P, src.data(), arena.data());
    std::vector<SeedParam> sp(P.lane_seed.size());
    for (size_t q = 0; q < sp.size(); ++q) {
        const int64_t si = P.lane_seed[q];
        if (si < 0) { sp[q] = SeedParam{ 0, 0, -1, 0, { -1, -1 }, { -1, -1 } }; continue; }
        const bsw_seed_task& s = tasks[si];
        sp[q] = SeedParam{ s.init_score, s.qbeg, s.h0 > 0 ? s.h0 : 0, s.id, { -1, -1 }, { -1, -1 } };
        if (gaps4) {
            sp[q].max_ins[0] = gaps4[4 * si + 0]; sp[q].max_del[0] = gaps4[4 * si + 1];
            sp[q].max_ins[1] = gaps4[4 * si + 2]; sp[q].max_del[1] = gaps4[4 * si + 3];
        }
    }
    std::vector<SeedRecord> rec(P.lane_seed.size());
    for (const Launch& L : P.launches) {
        if (L.kind != 5) return BSW_ERANGE;
        for (uint32_t t = L.tile0; t + 1 < L.tile0 + L.ntiles; t += 2) {
            const size_t base = (size_t)(t / 2) * K1_S;
#define EMU_CASE(V, G, S) if (variant == V && L.generic == G && sym == S) run_seed_pair<V, G, S>(dp, P2->w, P2->pen_clip5, P2->pen_clip3, P.tiles[t], P.tiles[t + 1], P.slots.data(), sp.data() + base, arena.data(), rec.data() + base, L.qmax, L.nqw_max);
            EMU_CASE(1, 0, 1) EMU_CASE(1, 0, 0) EMU_CASE(1, 1, 1) EMU_CASE(1, 1, 0)
            EMU_CASE(2, 0, 1) EMU_CASE(2, 0, 0) EMU_CASE(2, 1, 1) EMU_CASE(2, 1, 0)
#undef EMU_CASE
        }
    }
    for (size_t q = 0; q < rec.size(); ++q) {
        const int64_t si = P.lane_seed[q];
        if (si < 0) continue;
        const SeedRecord& r = rec[q];
        bsw_aln_record& o = out[si];
        o.id = r.id; o.qb = r.qb; o.qe = r.qe; o.rb = r.rb; o.re = r.re; o.score = r.score; o.truesc = r.truesc; o.w = r.w;
    }
    return BSW_OK;
}

// ---- host-logic self checks (tests/test_host_logic.py) ----
#include <sys/mman.h>
#include <unistd.h>
// Packs random sequences with both builds of bsw_pack.cpp and compares them with a scalar packer.  Every sequence ends
// flush against an inaccessible page, so a packer that reads past a sequence faults instead of passing.
// Returns 0, or 1000+i when task i differs, or a negative setup error.  *used_avx512 reports whether that build ran.
extern "C" int bsw_emu_pack_check(uint64_t seed, int ntasks, int max_len, int* used_avx512)
{
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t span = (((size_t)max_len + page - 1) / page + 1) * page;           // data pages + 1 guard page
    const bool wide = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("bmi2");
    if (used_avx512) *used_avx512 = wide ? 1 : 0;
    uint8_t* base = static_cast<uint8_t*>(mmap(nullptr, 2 * span * (size_t)ntasks, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
    if (base == MAP_FAILED) return -1;
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    std::vector<ExtTask> tasks((size_t)ntasks);
    for (int i = 0; i < ntasks; ++i) {
        for (int side = 0; side < 2; ++side) {
            uint8_t* blk = base + (size_t)(2 * i + side) * span;
            if (mprotect(blk + span - page, page, PROT_NONE)) return -2;
            const int len = 1 + (int)(rnd() % (uint64_t)max_len);
            uint8_t* sq = blk + span - page - (size_t)len;                          // last base = last readable byte
            const bool with_n = rnd() % 7 == 0;
            for (int k = 0; k < len; ++k) sq[k] = (uint8_t)((with_n && rnd() % 11 == 0) ? 4 : rnd() & 3);
            if (side == 0) { tasks[(size_t)i].q = sq; tasks[(size_t)i].qlen = len; } else { tasks[(size_t)i].t = sq; tasks[(size_t)i].tlen = len; }
        }
        tasks[(size_t)i].h0 = 1 + (int)(rnd() % 50); tasks[(size_t)i].w = (int)(rnd() % 100);
    }
    SchedOptions opt; opt.fast_matrix = true;
    const size_t bound = source_arena_bound(tasks.data(), (size_t)ntasks);
    int rc = 0;
    for (int isa = 0; isa < (wide ? 2 : 1) && !rc; ++isa) {
        std::vector<uint32_t> arena(bound, 0xdeadbeefu);
        std::vector<uint8_t> cls((size_t)ntasks); std::vector<SlotSrc> src((size_t)ntasks);
        size_t used = 0, bad = 0; std::string msg;
        const int e = (isa ? pack_tasks_avx512 : pack_tasks_sse2)(tasks.data(), (size_t)ntasks, 1, opt, cls.data(), src.data(), arena.data(), &used, &bad, &msg);
        if (e || used > bound) { rc = -3; break; }
        for (int i = 0; i < ntasks && !rc; ++i) {
            const ExtTask& t = tasks[(size_t)i];
            bool has_n = false;
            for (int side = 0; side < 2 && !rc; ++side) {
                const uint8_t* sq = side ? t.t : t.q; const int len = side ? t.tlen : t.qlen;
                const uint32_t* w = arena.data() + (size_t)(side ? src[(size_t)i].toff16 : src[(size_t)i].qoff16) * 4;
                const int words = ((len + 31) / 32) * 4;
                for (int k = 0; k < words * 8; ++k) {
                    const uint32_t got = (w[k >> 3] >> (4 * (k & 7))) & 15u, want = k < len ? sq[k] : 0u;
                    if (got != want) { rc = 1000 + i; break; }
                    if (want == 4) has_n = true;
                }
            }
            if (!rc && ((cls[(size_t)i] & 1) != (has_n ? 1 : 0))) rc = 1000 + i;
        }
        // a bad code anywhere must be rejected with the task's index
        if (!rc) {
            const int victim = (int)(rnd() % (uint64_t)ntasks);
            uint8_t* sq = const_cast<uint8_t*>(tasks[(size_t)victim].t);
            const int pos = (int)(rnd() % (uint64_t)tasks[(size_t)victim].tlen);
            const uint8_t keep = sq[pos]; sq[pos] = 5 + (uint8_t)(rnd() % 200);
            const int e2 = (isa ? pack_tasks_avx512 : pack_tasks_sse2)(tasks.data(), (size_t)ntasks, 1, opt, cls.data(), src.data(), arena.data(), &used, &bad, &msg);
            sq[pos] = keep;
            if (e2 != BSW_EINVAL || bad != (size_t)victim) rc = -4;
        }
    }
    munmap(base, 2 * span * (size_t)ntasks);
    return rc;
}

// BandClamp (tabulated) against clamp_band (the two divisions) over a grid of lengths, bands and penalties.
extern "C" int bsw_emu_band_clamp_check()
{
    int8_t mat[25];
    for (int a = 1; a <= 3; ++a)
        for (int oi = 0; oi <= 12; oi += 3)
            for (int ei = 1; ei <= 4; ++ei)
                for (int od = 0; od <= 12; od += 4)
                    for (int ed = 1; ed <= 3; ++ed)
                        for (int eb = 0; eb <= 10; eb += 5) {
                            for (int k = 0; k < 25; ++k) mat[k] = (int8_t)((k % 6 == 0 && k < 24) ? a : -4);
                            const BandClamp c(mat, eb, oi, ei, od, ed);
                            for (int q = 1; q <= 1400; q += (q < 300 ? 1 : 37))
                                for (int w = 0; w <= 400; w += 7)
                                    if (c(q, w) != clamp_band(mat, q, w, eb, oi, ei, od, ed)) return 1;
                        }
    return 0;
}
