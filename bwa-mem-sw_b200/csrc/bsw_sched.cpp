// bsw_sched.cpp -- length-bucketed scheduler + packer (see bsw_sched.h).
#include "bsw_sched.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <emmintrin.h>

#include "../../include/bsw.h"

namespace bsw {

int default_host_threads()
{
    unsigned n = std::thread::hardware_concurrency();
    if (n == 0) n = 1;
    if (n > 64) n = 64;
    return (int)n;
}

void parallel_for(size_t n, size_t grain, int nthreads, const void* ctx,
                  void (*fn)(const void* ctx, size_t lo, size_t hi))
{
    if (n == 0) return;
    if (nthreads <= 0) nthreads = default_host_threads();
    if (grain == 0) grain = 1;
    const size_t nchunks = (n + grain - 1) / grain;
    if ((size_t)nthreads > nchunks) nthreads = (int)nchunks;
    if (nthreads <= 1) { fn(ctx, 0, n); return; }
    std::atomic<size_t> next(0);
    auto body = [&]() {
        for (;;) {
            const size_t c = next.fetch_add(1);
            if (c >= nchunks) break;
            const size_t lo = c * grain, hi = std::min(n, lo + grain);
            fn(ctx, lo, hi);
        }
    };
    std::vector<std::thread> th;
    th.reserve((size_t)nthreads - 1);
    for (int k = 1; k < nthreads; ++k) th.emplace_back(body);
    body();
    for (auto& x : th) x.join();
}

int clamp_band(const int8_t mat[25], int qlen, int w, int end_bonus, int o_ins, int e_ins, int o_del, int e_del)
{
    int mx = 0;
    for (int k = 0; k < 25; ++k) mx = mx > mat[k] ? mx : mat[k];
    int max_ins = (int)((double)(qlen * mx + end_bonus - o_ins) / e_ins + 1.);
    max_ins = max_ins > 1 ? max_ins : 1;
    w = w < max_ins ? w : max_ins;
    int max_del = (int)((double)(qlen * mx + end_bonus - o_del) / e_del + 1.);
    max_del = max_del > 1 ? max_del : 1;
    w = w < max_del ? w : max_del;
    return w;
}

size_t source_arena_bound(const ExtTask* tasks, size_t n)
{
    size_t words = 0;
    for (size_t i = 0; i < n; ++i) {
        const size_t ql = tasks[i].qlen > 0 ? (size_t)tasks[i].qlen : 0, tl = tasks[i].tlen > 0 ? (size_t)tasks[i].tlen : 0;
        words += ((ql + 31) / 32 + (tl + 31) / 32) * 4;
    }
    return words + 64;
}

int pack_tasks(const ExtTask* tasks, size_t n, int max_mat, const SchedOptions& opt, uint8_t* cls, SlotSrc* src,
               uint32_t* arena, size_t* words_used, size_t* bad_task, std::string* msg)
{
    // bsw_pack.cpp is compiled twice; the AVX-512 build packs 64 bases per step with masked loads (no tail branches)
    static const bool wide = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
                             __builtin_cpu_supports("bmi2") && !getenv("BSW_NO_AVX512");
    return wide ? pack_tasks_avx512(tasks, n, max_mat, opt, cls, src, arena, words_used, bad_task, msg)
                : pack_tasks_sse2(tasks, n, max_mat, opt, cls, src, arena, words_used, bad_task, msg);
}

int64_t pack2_flat(const uint8_t* qbuf, const int64_t* qoff, const uint8_t* tbuf, const int64_t* toff, size_t count,
                   uint32_t* arena, SlotSrc* src, std::vector<uint32_t>* n_list)
{
    static const bool wide = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
                             __builtin_cpu_supports("bmi2") && !getenv("BSW_NO_AVX512");
    return wide ? pack2_flat_avx512(qbuf, qoff, tbuf, toff, count, arena, src, n_list)
                : pack2_flat_generic(qbuf, qoff, tbuf, toff, count, arena, src, n_list);
}

// order[] = task indices by ascending 32-bit key (stable LSD radix).  The digit width follows the chunk size: three
// 11-bit passes for the usual 8-32 k task chunk (the three histograms come from one pass over the keys and stay in
// L1), two 16-bit passes for big plans (resident batches) where the per-pass traffic dominates.
static void radix_order(const uint32_t* key, size_t n, std::vector<uint32_t>& order, std::vector<uint32_t>& tmp, std::vector<uint32_t>& hist,
                        bool key22 = false)
{
    if (key22) {                                       // keys below 2^22: two 11-bit passes
        hist.assign(2 * 2049, 0);
        uint32_t* h0 = hist.data(); uint32_t* h1 = h0 + 2049;
        for (size_t i = 0; i < n; ++i) { const uint32_t k = key[i]; ++h0[(k & 2047u) + 1]; ++h1[(k >> 11) + 1]; }
        for (size_t b = 0; b < 2048; ++b) { h0[b + 1] += h0[b]; h1[b + 1] += h1[b]; }
        for (size_t i = 0; i < n; ++i) tmp[h0[key[i] & 2047u]++] = (uint32_t)i;
        for (size_t i = 0; i < n; ++i) { const uint32_t t = tmp[i]; order[h1[key[t] >> 11]++] = t; }
        return;
    }
    if (n > ((size_t)1 << 17)) {
        hist.assign(2 * 65537, 0);
        uint32_t* h0 = hist.data(); uint32_t* h1 = h0 + 65537;
        for (size_t i = 0; i < n; ++i) { ++h0[(key[i] & 0xffffu) + 1]; ++h1[(key[i] >> 16) + 1]; }
        for (size_t b = 0; b < 65536; ++b) { h0[b + 1] += h0[b]; h1[b + 1] += h1[b]; }
        for (size_t i = 0; i < n; ++i) tmp[h0[key[i] & 0xffffu]++] = (uint32_t)i;
        for (size_t i = 0; i < n; ++i) { const uint32_t t = tmp[i]; order[h1[key[t] >> 16]++] = t; }
        return;
    }
    hist.assign(3 * 2049, 0);
    uint32_t* h0 = hist.data(); uint32_t* h1 = h0 + 2049; uint32_t* h2 = h1 + 2049;
    for (size_t i = 0; i < n; ++i) { const uint32_t k = key[i]; ++h0[(k & 2047u) + 1]; ++h1[((k >> 11) & 2047u) + 1]; ++h2[(k >> 22) + 1]; }
    for (size_t b = 0; b < 2048; ++b) { h0[b + 1] += h0[b]; h1[b + 1] += h1[b]; h2[b + 1] += h2[b]; }
    for (size_t i = 0; i < n; ++i) order[h0[key[i] & 2047u]++] = (uint32_t)i;
    for (size_t i = 0; i < n; ++i) { const uint32_t t = order[i]; tmp[h1[(key[t] >> 11) & 2047u]++] = t; }
    for (size_t i = 0; i < n; ++i) { const uint32_t t = tmp[i]; order[h2[key[t] >> 22]++] = t; }
}

static inline size_t k1_tile_smem(int qmax, int nqw)
{
    return 128 + ((size_t)(nqw + 8) + (size_t)(qmax + 1 + K1_EH_SLACK)) * TILE_LANES * 4u;
}
static inline size_t k2_task_smem(int qmax, int wmax)
{
    const size_t qcap = ((size_t)qmax + 1 + 255) & ~(size_t)255;
    const bool ring = qcap > 2048 && 2 * (size_t)wmax + 513 <= 2048;
    const size_t rcap = ring ? 2048 : qcap;
    return 128 + ((ring ? 0 : (qcap >> 3)) + 4 + (rcap >> 5) + 4 + rcap + 16) * 4u + 16u;
}
static inline int occupancy(size_t smem)
{
    const size_t budget = 227 * 1024;
    size_t n = budget / (smem + 1024);          // 1 KB per CTA reserved by the driver
    if (n > 32) n = 32;
    if (n < 1) n = 1;
    return (int)n;
}

void build_plan(const ExtTask* tasks, const uint8_t* cls, const SlotSrc* src, size_t n, const SchedOptions& opt, Plan* plan)
{
    (void)opt;
    plan->tiles.clear(); plan->slots.clear(); plan->slot_src.clear(); plan->slot_task.clear(); plan->launches.clear();
    plan->tiled_words = 0; plan->est_cells = 0; plan->n_k1_tiles = 0;
    if (n == 0) return;

    // Sort key (32 bit, descending cost first).  The order is only a scheduling heuristic (any order gives the same
    // results), but it decides how much of a tile's time its lanes idle: the lanes advance row by row together, so a tile
    // costs (rows of its longest lane) x (width of its widest live window).  The window of a row is about min(qlen, h0)
    // wide (the non-zero run around the diagonal grows with the score budget): tlen * min(qlen, h0) predicts the cell
    // count of a 150 bp task with correlation 0.99 (tools/tile_efficiency.py).
    //   chunk plans (< 100 k tasks), K1 classes: [class:3][qlen/16:7][min(qlen,h0)/4:6][tlen/8:6] (22 bits) -- a coarse qlen
    //       bucket (it sets the tile's shared memory, i.e. the occupancy bucket), then the window width, then the rows.
    //       Lane efficiency of 16 k task chunks (row-lockstep model) against the key below: 150 bp reads 0.89 -> 0.96,
    //       50-250 bp mix 0.67 -> 0.87; measured e2e on 1 M tasks 6.45 -> 5.85 ms and 11.9 -> 8.3 ms.
    //   big plans and K2 classes: [class:3][qlen:13][tlen/4:10][h0/2:6] (K2: qlen/8) -- with thousands of tasks per exact
    //       qlen the three-level sort is already 0.99 efficient, and it measured 5 % faster there than the width key
    //       (1 107 vs 1 050 GCUPS on the resident 1 M x 150 bp plan).
    // classes: 0 K1 fast, 1 K1 matrix, 4 K2 fast, 5 K2 matrix
    std::vector<uint32_t>& key = plan->key; std::vector<uint32_t>& order = plan->order;
    std::vector<uint32_t>& tmp = plan->tmp; std::vector<uint32_t>& hist = plan->hist;
    key.resize(n); order.resize(n); tmp.resize(n);
    const bool width_key = n < 100000;
    if (width_key) {
        // 22-bit keys (two radix passes): [class:3][qlen/16:7][min(qlen,h0)/4:6][tlen/8:6]; K2 classes [class:3][qlen/128:7][tlen/512:12]
        for (size_t i = 0; i < n; ++i) {
            const ExtTask& t = tasks[i];
            const uint32_t c = ((cls[i] & 2u) ? 4u : 0u) | (cls[i] & 1u);
            if (cls[i] & 2u) {
                const uint32_t ql = (uint32_t)std::min(t.qlen >> 7, 127), tl = (uint32_t)std::min(t.tlen >> 9, 4095);
                key[i] = (c << 19) | ((127u - ql) << 12) | (4095u - tl);
            } else {
                const uint32_t qb = (uint32_t)std::min(t.qlen >> 4, 127);
                const uint32_t wd = (uint32_t)std::min(std::min(t.qlen, t.h0) >> 2, 63), tl = (uint32_t)std::min(t.tlen >> 3, 63);
                key[i] = (c << 19) | ((127u - qb) << 12) | ((63u - wd) << 6) | (63u - tl);
            }
        }
    } else {
        for (size_t i = 0; i < n; ++i) {
            const ExtTask& t = tasks[i];
            const uint32_t c = ((cls[i] & 2u) ? 4u : 0u) | (cls[i] & 1u);
            const uint32_t ql = (cls[i] & 2u) ? (uint32_t)std::min(t.qlen >> 3, 8191) : (uint32_t)std::min(t.qlen, 8191);
            const uint32_t tl = (uint32_t)std::min(t.tlen >> 2, 1023), h = (uint32_t)std::min(t.h0 >> 1, 63);
            key[i] = (c << 29) | ((8191u - ql) << 16) | ((1023u - tl) << 6) | (63u - h);
        }
    }
    const int class_shift = width_key ? 19 : 29;
    radix_order(key.data(), n, order, tmp, hist, width_key);

    // new occupancy bucket (= new launch) when a tile would fit at >= 1.15x the CTAs/SM of the current bucket (big plans;
    // measured best with the launches spread over four streams), >= 1.3x (100-400 k tasks) or >= 2x (the chunks of the
    // host pipeline: every extra launch and stream costs more there than the occupancy it buys -- e2e 8.0 -> 6.5 ms on
    // 1 M x 150 bp)
    static const int bucket_env = getenv("BSW_BUCKET_PCT") ? atoi(getenv("BSW_BUCKET_PCT")) : 0;
    const int bucket_pct = bucket_env ? bucket_env : (n >= 400000 ? 115 : (n >= 100000 ? 130 : 200));
    // tiles
    const size_t nslot_bound = n + (size_t)8 * TILE_LANES;
    plan->slots.reserve(nslot_bound); plan->slot_src.reserve(nslot_bound); plan->slot_task.reserve(nslot_bound);
    plan->tiles.reserve(n / TILE_LANES + 16);
    size_t arena16 = 0;     // tiled arena, in 16-byte units
    auto close_launch = [&](Launch& L, uint32_t tile_end) {
        L.ntiles = tile_end - L.tile0;
        if (L.ntiles) plan->launches.push_back(L);
    };
    size_t i = 0;
    while (i < n) {
        const uint32_t c = key[order[i]] >> class_shift;
        size_t cend = i;
        while (cend < n && (key[order[cend]] >> class_shift) == c) ++cend;
        const bool is_k2 = c >= 4u;
        Launch L{};
        L.kind = is_k2 ? 2 : 1; L.generic = (int)(c & 1u); L.tile0 = (uint32_t)plan->tiles.size();
        int occ0 = 0;
        while (i < cend) {
            // one K2 task or one K1 tile of 32 tasks
            const size_t ntask = is_k2 ? 1 : std::min<size_t>(TILE_LANES, cend - i);
            const int nsub = 1;
            TileHdr hd[2];
            int qmax = 0, nqw_max = 0, wmax = 0;
            for (int sub = 0; sub < nsub; ++sub) {
                hd[sub] = TileHdr{};
                hd[sub].slot0 = (uint32_t)plan->slots.size();
                int tq = 0, tt = 0;
                const size_t lanes = is_k2 ? 1 : (size_t)TILE_LANES;
                if (ntask == lanes) {
                    // a full tile (the common case): write the 32 slots through raw pointers
                    const size_t s0 = plan->slots.size();
                    plan->slots.resize(s0 + lanes); plan->slot_src.resize(s0 + lanes); plan->slot_task.resize(s0 + lanes);
                    SlotParam* sp = plan->slots.data() + s0; SlotSrc* ss = plan->slot_src.data() + s0; int64_t* st = plan->slot_task.data() + s0;
                    const uint32_t* ord = order.data() + i;
                    uint64_t cells = 0;
                    for (size_t l = 0; l < lanes; ++l) {
                        const uint32_t ti = ord[l];
                        const ExtTask& t = tasks[ti];
                        tq = std::max(tq, t.qlen); tt = std::max(tt, t.tlen); wmax = std::max(wmax, t.w);
                        sp[l] = SlotParam{ t.qlen, t.tlen, t.h0, t.w };
                        ss[l] = src[ti];
                        st[l] = (int64_t)ti;
                        cells += (uint64_t)(std::min<int64_t>(t.qlen, 2 * (int64_t)t.w + 1) * t.tlen);
                    }
                    plan->est_cells += cells;
                } else
                for (size_t l = 0; l < lanes; ++l) {
                    const size_t k = l;
                    if (k < ntask) {
                        const uint32_t ti = order[i + k];
                        const ExtTask& t = tasks[ti];
                        tq = std::max(tq, t.qlen); tt = std::max(tt, t.tlen); wmax = std::max(wmax, t.w);
                        plan->slots.push_back(SlotParam{ t.qlen, t.tlen, t.h0, t.w });
                        plan->slot_src.push_back(src[ti]);
                        plan->slot_task.push_back((int64_t)ti);
                        const int64_t band = std::min<int64_t>(t.qlen, 2 * (int64_t)t.w + 1);
                        plan->est_cells += (uint64_t)(band * t.tlen);
                    } else {
                        plan->slots.push_back(SlotParam{ 0, 0, 0, 0 });
                        plan->slot_src.push_back(SlotSrc{ 0, 0 });
                        plan->slot_task.push_back(-1);
                    }
                }
                const int nqw = (tq + 7) >> 3;
                const int ntw = (tt + 7) >> 3;
                if (is_k2) {
                    hd[sub].qoff16 = src[order[i]].qoff16;                      // K2 reads the source arena directly
                    hd[sub].toff16 = src[order[i]].toff16;
                } else {
                    hd[sub].qoff16 = (uint32_t)arena16; arena16 += (size_t)nqw * TILE_LANES * 4 / 16;
                    hd[sub].toff16 = (uint32_t)arena16; arena16 += (size_t)ntw * TILE_LANES * 4 / 16;
                    ++plan->n_k1_tiles;
                }
                hd[sub].nqw_ntw = (uint32_t)nqw | ((uint32_t)ntw << 16);
                qmax = std::max(qmax, tq); nqw_max = std::max(nqw_max, nqw);
            }
            const size_t smem = is_k2 ? k2_task_smem(qmax, wmax) : k1_tile_smem(qmax, nqw_max);
            // bucket boundary: start a new launch when this tile would fit at >= 1.3x the occupancy of the launch
            const int occ = occupancy(smem);
            L.wmax = std::max(L.wmax, wmax);
            if (occ0 == 0) { occ0 = occ; L.qmax = qmax; L.nqw_max = nqw_max; }
            else if (qmax > L.qmax) { L.qmax = qmax; L.nqw_max = std::max(L.nqw_max, nqw_max); }   // saturated sort key (very long tasks)
            else if (!is_k2 && occ * 100 >= occ0 * bucket_pct) {      // K2: one launch per class (measured: bucket tails cost more than occupancy gains)
                close_launch(L, (uint32_t)plan->tiles.size());
                L.tile0 = (uint32_t)plan->tiles.size(); L.qmax = qmax; L.nqw_max = nqw_max; L.wmax = wmax; occ0 = occ;
            }
            for (int sub = 0; sub < nsub; ++sub) plan->tiles.push_back(hd[sub]);
            i += ntask;
        }
        close_launch(L, (uint32_t)plan->tiles.size());
    }
    plan->tiled_words = arena16 * 4 + 32;
}

bool build_dp_plan(const ExtTask* tasks, const uint8_t* cls, size_t n, const SchedOptions& opt, Plan* plan, DpGeometry* g)
{
    if (n == 0) return false;
    DpBuckets bks;
    memset(&bks, 0, sizeof(bks));
    for (size_t i = 0; i < n; ++i) {
        if (cls[i] & 6u) return false;                                   // a long task: the host planner handles the chunk
        bks.add(cls[i] & 1, tasks[i].qlen, tasks[i].tlen, tasks[i].w);
    }
    return dp_geometry(bks, n, opt, plan, g);
}

bool dp_geometry(const DpBuckets& bks, size_t n, const SchedOptions& opt, Plan* plan, DpGeometry* g)
{
    (void)opt;
    plan->tiles.clear(); plan->slots.clear(); plan->slot_src.clear(); plan->slot_task.clear(); plan->launches.clear();
    plan->tiled_words = 0; plan->est_cells = 0; plan->n_k1_tiles = 0;
    if (n == 0) return false;
    const DpBuckets::B (*bk)[128] = bks.bk;
    static const int bucket_env = getenv("BSW_BUCKET_PCT") ? atoi(getenv("BSW_BUCKET_PCT")) : 0;
    const int bucket_pct = bucket_env ? bucket_env : (n >= 400000 ? 115 : (n >= 100000 ? 130 : 200));
    size_t arena16 = 0;
    uint32_t slot = 0, pos = 0;
    g->nmajor = 0;
    memset(g->major_of, 0, sizeof(g->major_of));
    for (int c = 0; c < 2; ++c) {
        uint32_t cnt = 0;
        for (int b = 127; b >= 0; --b) {
            if (!bk[c][b].cnt) continue;
            if (g->nmajor >= 192) return false;
            g->major_of[(c << 7) | b] = (uint8_t)g->nmajor;
            g->major_start[g->nmajor++] = pos + cnt;
            cnt += bk[c][b].cnt;
        }
        g->class_count[c] = cnt; g->class_pos0[c] = pos; g->class_slot0[c] = slot; g->class_tile0[c] = (uint32_t)plan->tiles.size();
        pos += cnt;
        if (!cnt) continue;
        Launch L{};
        L.kind = 1; L.generic = c; L.tile0 = (uint32_t)plan->tiles.size();
        int occ0 = 0;
        int b = 127;                                                     // buckets in sorted order: longest first
        uint32_t left = bk[c][b].cnt;
        for (uint32_t done = 0; done < cnt; done += TILE_LANES) {
            int need = (int)std::min<uint32_t>(TILE_LANES, cnt - done);
            int qmax = 0, tmax = 0, wmax = 0;
            while (need > 0) {
                while (left == 0) { --b; left = bk[c][b].cnt; }
                const uint32_t take = std::min<uint32_t>(left, (uint32_t)need);
                qmax = std::max(qmax, bk[c][b].maxq); tmax = std::max(tmax, bk[c][b].maxt); wmax = std::max(wmax, bk[c][b].maxw);
                left -= take; need -= (int)take;
            }
            const int nqw = (qmax + 7) >> 3, ntw = (tmax + 7) >> 3;
            TileHdr hd{};
            hd.slot0 = slot; slot += TILE_LANES;
            hd.qoff16 = (uint32_t)arena16; arena16 += (size_t)nqw * TILE_LANES * 4 / 16;
            hd.toff16 = (uint32_t)arena16; arena16 += (size_t)ntw * TILE_LANES * 4 / 16;
            hd.nqw_ntw = (uint32_t)nqw | ((uint32_t)ntw << 16);
            const int occ = occupancy(k1_tile_smem(qmax, nqw));
            L.wmax = std::max(L.wmax, wmax);
            if (occ0 == 0) { occ0 = occ; L.qmax = qmax; L.nqw_max = nqw; }
            else if (qmax > L.qmax) { L.qmax = qmax; L.nqw_max = std::max(L.nqw_max, nqw); }
            else if (occ * 100 >= occ0 * bucket_pct) {
                L.ntiles = (uint32_t)plan->tiles.size() - L.tile0;
                if (L.ntiles) plan->launches.push_back(L);
                L.tile0 = (uint32_t)plan->tiles.size(); L.qmax = qmax; L.nqw_max = nqw; L.wmax = wmax; occ0 = occ;
            }
            plan->tiles.push_back(hd);
            ++plan->n_k1_tiles;
        }
        L.ntiles = (uint32_t)plan->tiles.size() - L.tile0;
        if (L.ntiles) plan->launches.push_back(L);
    }
    g->ntiles = (uint32_t)plan->tiles.size();
    g->nslots = slot;
    plan->tiled_words = arena16 * 4 + 32;
    return true;
}

static inline size_t k3_tile_smem(int qmax, int nqw)
{
    return 128 + ((size_t)2 * (size_t)(nqw + 8) + (size_t)(qmax + 1 + K1_EH_SLACK)) * TILE_LANES * 4u;
}

void build_seed_plan(const ExtTask* tasks, const uint8_t* cls, const SlotSrc* src, size_t nseeds, const SchedOptions& opt, Plan* plan)
{
    (void)opt;
    plan->tiles.clear(); plan->slots.clear(); plan->slot_src.clear(); plan->slot_task.clear(); plan->launches.clear();
    plan->lane_seed.clear();
    plan->tiled_words = 0; plan->est_cells = 0; plan->n_k1_tiles = 0;
    if (nseeds == 0) return;
    const size_t n = nseeds;
    std::vector<uint32_t>& key = plan->key; std::vector<uint32_t>& order = plan->order;
    std::vector<uint32_t>& tmp = plan->tmp; std::vector<uint32_t>& hist = plan->hist;
    key.resize(n); order.resize(n); tmp.resize(n);
    // Sort key: the lanes of a K3 tile run the left extensions together, then the right ones, so a tile costs (longest
    // left flank) + (longest right flank) and both must be alike across its lanes.  [matrix class:1][longer flank/16:7
    // (shared memory of the tile pair)][estimated left cells/512:12][estimated right cells/128:12], cells estimated as
    // tlen * min(qlen, score budget) like build_plan does (the right flank starts from about h0 + left qlen, which the
    // caller passes in the right task's otherwise unused w).  Modelled
    // lane efficiency of 8 k seed chunks (tools/seed_tile_efficiency.py): 0.52 with the former (longer flank, mean flank,
    // h0) key -> 0.84.
    for (size_t i = 0; i < n; ++i) {
        const ExtTask& l = tasks[2 * i]; const ExtTask& r = tasks[2 * i + 1];
        const uint32_t g = ((cls[2 * i] | cls[2 * i + 1]) & 1u);
        const uint32_t qb = (uint32_t)std::min(std::max(l.qlen, r.qlen) >> 4, 127);
        const int64_t el = l.qlen > 0 ? (int64_t)l.tlen * std::min(l.qlen, l.h0) : 0;
        const int64_t er = r.qlen > 0 ? (int64_t)r.tlen * std::min(r.qlen, std::max(r.w, 1)) : 0;     // r.w: budget hint set by the caller
        const uint32_t bl = (uint32_t)std::min<int64_t>(el >> 9, 4095), br = (uint32_t)std::min<int64_t>(er >> 7, 4095);
        key[i] = (g << 31) | ((127u - qb) << 24) | ((4095u - bl) << 12) | (4095u - br);
    }
    radix_order(key.data(), n, order, tmp, hist);

    size_t arena16 = 0;
    auto close_launch = [&](Launch& L, uint32_t tile_end) {
        L.ntiles = tile_end - L.tile0;
        if (L.ntiles) plan->launches.push_back(L);
    };
    size_t i = 0;
    while (i < n) {
        const uint32_t g = key[order[i]] >> 31;
        size_t cend = i;
        while (cend < n && (key[order[cend]] >> 31) == g) ++cend;
        Launch L{};
        L.kind = 5; L.generic = (int)g; L.tile0 = (uint32_t)plan->tiles.size();
        int occ0 = 0;
        while (i < cend) {
            const size_t ns = std::min<size_t>(TILE_LANES, cend - i);
            TileHdr hd[2];
            int qmax = 0, nqw_max = 0;
            for (int side = 0; side < 2; ++side) {
                hd[side] = TileHdr{};
                hd[side].slot0 = (uint32_t)plan->slots.size();
                int tq = 0, tt = 0;
                for (size_t l = 0; l < (size_t)TILE_LANES; ++l) {
                    if (l < ns) {
                        const size_t ti = 2 * (size_t)order[i + l] + (size_t)side;
                        const ExtTask& t = tasks[ti];
                        tq = std::max(tq, t.qlen); tt = std::max(tt, t.qlen > 0 ? t.tlen : 0);
                        plan->slots.push_back(SlotParam{ t.qlen, t.qlen > 0 ? t.tlen : 0, t.h0, 0 });
                        plan->slot_src.push_back(src[ti]);
                        plan->slot_task.push_back((int64_t)ti);
                        plan->est_cells += (uint64_t)t.qlen * (uint64_t)(t.qlen > 0 ? t.tlen : 0);
                        if (side == 0) plan->lane_seed.push_back((int64_t)order[i + l]);
                    } else {
                        plan->slots.push_back(SlotParam{ 0, 0, 0, 0 });
                        plan->slot_src.push_back(SlotSrc{ 0, 0 });
                        plan->slot_task.push_back(-1);
                        if (side == 0) plan->lane_seed.push_back(-1);
                    }
                }
                const int nqw = (tq + 7) >> 3, ntw = (tt + 7) >> 3;
                hd[side].qoff16 = (uint32_t)arena16; arena16 += (size_t)nqw * TILE_LANES * 4 / 16;
                hd[side].toff16 = (uint32_t)arena16; arena16 += (size_t)ntw * TILE_LANES * 4 / 16;
                hd[side].nqw_ntw = (uint32_t)nqw | ((uint32_t)ntw << 16);
                ++plan->n_k1_tiles;
                qmax = std::max(qmax, tq); nqw_max = std::max(nqw_max, nqw);
            }
            const int occ = occupancy(k3_tile_smem(qmax, nqw_max));
            if (occ0 == 0) { occ0 = occ; L.qmax = qmax; L.nqw_max = nqw_max; }
            else if (qmax > L.qmax) { L.qmax = qmax; L.nqw_max = std::max(L.nqw_max, nqw_max); }
            else if (occ * 10 >= occ0 * 13) {
                close_launch(L, (uint32_t)plan->tiles.size());
                L.tile0 = (uint32_t)plan->tiles.size(); L.qmax = qmax; L.nqw_max = nqw_max; occ0 = occ;
            }
            plan->tiles.push_back(hd[0]); plan->tiles.push_back(hd[1]);
            i += ns;
        }
        close_launch(L, (uint32_t)plan->tiles.size());
    }
    plan->tiled_words = arena16 * 4 + 32;
}

}  // namespace bsw
