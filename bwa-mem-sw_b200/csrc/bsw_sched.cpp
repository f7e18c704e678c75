// bsw_sched.cpp -- length-bucketed scheduler + packer (see bsw_sched.h).
#include "bsw_sched.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

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

// ---- base scanning: 8 bases per 64-bit word ----
// returns bit0 = some base is N (4), bit1 = some base code > 4 (invalid)
static inline unsigned scan_bases(const uint8_t* s, int n)
{
    uint64_t any4 = 0, bad = 0;
    int k = 0;
    for (; k + 8 <= n; k += 8) {
        uint64_t x;
        memcpy(&x, s + k, 8);
        any4 |= x & 0x0404040404040404ull;
        bad |= (x | (x + 0x7b7b7b7b7b7b7b7bull)) & 0x8080808080808080ull;      // byte >= 5
    }
    for (; k < n; ++k) { any4 |= (uint64_t)(s[k] & 4u); bad |= (uint64_t)(s[k] > 4u ? 0x80u : 0u); }
    return (any4 ? 1u : 0u) | (bad ? 2u : 0u);
}

int validate_tasks(const ExtTask* tasks, size_t n, int max_mat, const SchedOptions& opt,
                   uint8_t* cls, size_t* bad_task, std::string* msg)
{
    std::atomic<int> err(0);
    std::atomic<size_t> bad(n);
    const int k1cap = K1_QLEN_CAP;
    pfor(n, 4096, opt.host_threads, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            const ExtTask& t = tasks[i];
            int e = 0;
            if (!t.q || !t.t || t.qlen < 1 || t.tlen < 1 || t.h0 < 1 || t.w < 0) e = BSW_EINVAL;
            else if ((int64_t)t.h0 + (int64_t)t.qlen * max_mat > SCORE_CAP || t.qlen > K2_QLEN_CAP || t.tlen > 500000) e = BSW_ERANGE;
            unsigned f = 0;
            if (!e) {
                f = scan_bases(t.q, t.qlen) | scan_bases(t.t, t.tlen);
                if (f & 2u) e = BSW_EINVAL;
            }
            uint8_t c = 0;
            if (!e) {
                if ((f & 1u) || !opt.fast_matrix) c |= 1;
                bool longtask = opt.force_kernel == 2 || (opt.force_kernel == 0 && t.qlen >= opt.k2_min_qlen) || t.qlen > k1cap;
                if (opt.variant == 2) {               // K2 implements the V1 recurrence only
                    longtask = false;
                    if (t.qlen > k1cap || opt.force_kernel == 2) e = BSW_ERANGE;
                }
                if (longtask) c |= 2;
            }
            cls[i] = c;
            if (e) {
                int expect = 0;
                err.compare_exchange_strong(expect, e);
                size_t cur = bad.load();
                while (i < cur && !bad.compare_exchange_weak(cur, i)) {}
            }
        }
    });
    if (err.load()) {
        if (bad_task) *bad_task = bad.load();
        if (msg) {
            const size_t b = bad.load();
            const ExtTask& t = tasks[b];
            *msg = "task " + std::to_string(b) + ": qlen=" + std::to_string(t.qlen) + " tlen=" + std::to_string(t.tlen) +
                   " h0=" + std::to_string(t.h0) + " w=" + std::to_string(t.w) +
                   (err.load() == BSW_ERANGE ? " outside the numeric envelope (16-bit row state / length caps / V2 long task)"
                                             : " invalid (null pointer, length < 1, h0 < 1 or base code > 4)");
        }
        return err.load();
    }
    return 0;
}

static inline size_t k1_tile_smem(int qmax, int nqw)
{
    return 128 + ((size_t)(nqw + 1) + (size_t)(qmax + 1 + K1_EH_SLACK)) * TILE_LANES * 4u;
}
static inline size_t k2_task_smem(int qmax)
{
    const size_t qcap = ((size_t)qmax + 1 + 255) & ~(size_t)255;
    return 128 + ((qcap >> 3) + 4 + (qcap >> 5) + 4 + qcap + 8) * 4u + 16u;
}
static inline int occupancy(size_t smem)
{
    const size_t budget = 227 * 1024;
    size_t n = budget / (smem + 1024);          // 1 KB per CTA reserved by the driver
    if (n > 32) n = 32;
    if (n < 1) n = 1;
    return (int)n;
}

void build_plan(const ExtTask* tasks, const uint8_t* cls, size_t n, const SchedOptions& opt, Plan* plan)
{
    plan->tiles.clear(); plan->slots.clear(); plan->slot_task.clear(); plan->launches.clear();
    plan->arena_words = 0; plan->est_cells = 0;
    if (n == 0) return;

    // sort key: class (K1 fast, K1 generic, K2 fast, K2 generic), then longest query first, then target, then h0
    std::vector<uint64_t> key(n);
    std::vector<uint32_t> order(n);
    pfor(n, 16384, opt.host_threads, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            const ExtTask& t = tasks[i];
            const uint64_t c = ((cls[i] & 2u) ? 2u : 0u) | (cls[i] & 1u);
            const uint64_t q = 0xffffu - (uint64_t)t.qlen, tl = 0xfffffu - (uint64_t)t.tlen;
            const uint64_t h = 0xffffu - (uint64_t)(t.h0 > 0xffff ? 0xffff : t.h0);
            key[i] = (c << 52) | (q << 36) | (tl << 16) | h;
            order[i] = (uint32_t)i;
        }
    });
    // bucket by (class, qlen) with a counting pass, then sort inside the buckets in parallel
    {
        const size_t NB = 4u << 16;
        std::vector<uint32_t> cnt(NB + 1, 0);
        for (size_t i = 0; i < n; ++i) ++cnt[(size_t)(key[i] >> 36) + 1];
        for (size_t b = 0; b < NB; ++b) cnt[b + 1] += cnt[b];
        std::vector<uint32_t> pos(cnt.begin(), cnt.end() - 1);
        for (size_t i = 0; i < n; ++i) order[pos[(size_t)(key[i] >> 36)]++] = (uint32_t)i;
        std::vector<std::pair<uint32_t, uint32_t>> ranges;
        for (size_t b = 0; b < NB; ++b) if (cnt[b + 1] - cnt[b] > 1) ranges.emplace_back(cnt[b], cnt[b + 1]);
        pfor(ranges.size(), 1, opt.host_threads, [&](size_t lo, size_t hi) {
            for (size_t r = lo; r < hi; ++r)
                std::sort(order.begin() + ranges[r].first, order.begin() + ranges[r].second,
                          [&](uint32_t a, uint32_t b) { return key[a] != key[b] ? key[a] < key[b] : a < b; });
        });
    }

    // tiles
    size_t arena16 = 0;     // in 16-byte units
    auto close_launch = [&](Launch& L, uint32_t tile_end) {
        L.ntiles = tile_end - L.tile0;
        if (L.ntiles) plan->launches.push_back(L);
    };
    size_t i = 0;
    while (i < n) {
        const uint32_t c = (uint32_t)(key[order[i]] >> 52);
        size_t cend = i;
        while (cend < n && (uint32_t)(key[order[cend]] >> 52) == c) ++cend;
        const bool is_k2 = (c & 2u) != 0;
        Launch L{};
        L.kind = is_k2 ? 2 : 1; L.generic = (int)(c & 1u); L.tile0 = (uint32_t)plan->tiles.size();
        int occ0 = 0;
        while (i < cend) {
            TileHdr hd{};
            hd.slot0 = (uint32_t)plan->slots.size();
            int qmax = 0, tmax = 0;
            const size_t ntask = is_k2 ? 1 : std::min<size_t>(TILE_LANES, cend - i);
            for (size_t k = 0; k < ntask; ++k) {
                const ExtTask& t = tasks[order[i + k]];
                qmax = std::max(qmax, t.qlen); tmax = std::max(tmax, t.tlen);
                plan->slots.push_back(SlotParam{ t.qlen, t.tlen, t.h0, t.w });
                plan->slot_task.push_back((int64_t)order[i + k]);
                const int64_t band = std::min<int64_t>(t.qlen, 2 * (int64_t)t.w + 1);
                plan->est_cells += (uint64_t)(band * t.tlen);
            }
            if (!is_k2)
                for (size_t k = ntask; k < (size_t)TILE_LANES; ++k) {
                    plan->slots.push_back(SlotParam{ 0, 0, 0, 0 });
                    plan->slot_task.push_back(-1);
                }
            const int nqw = (qmax + 7) >> 3, ntw = (tmax + 7) >> 3;
            size_t smem;
            if (is_k2) {
                hd.qoff16 = (uint32_t)arena16; arena16 += ((size_t)nqw * 4 + 15) / 16;
                hd.toff16 = (uint32_t)arena16; arena16 += ((size_t)ntw * 4 + 15) / 16;
                arena16 = (arena16 + 7) & ~(size_t)7;                     // 128-byte alignment of the next block
                smem = k2_task_smem(qmax);
            } else {
                hd.qoff16 = (uint32_t)arena16; arena16 += (size_t)nqw * TILE_LANES * 4 / 16;
                hd.toff16 = (uint32_t)arena16; arena16 += (size_t)ntw * TILE_LANES * 4 / 16;
                smem = k1_tile_smem(qmax, nqw);
            }
            hd.nqw_ntw = (uint32_t)nqw | ((uint32_t)ntw << 16);
            // bucket boundary: start a new launch when this tile would fit at >= 1.3x the occupancy of the launch
            const int occ = occupancy(smem);
            if (occ0 == 0) { occ0 = occ; L.qmax = qmax; L.nqw_max = nqw; }
            else if (occ * 10 >= occ0 * 13) {
                close_launch(L, (uint32_t)plan->tiles.size());
                L.tile0 = (uint32_t)plan->tiles.size(); L.qmax = qmax; L.nqw_max = nqw; occ0 = occ;
            }
            plan->tiles.push_back(hd);
            i += ntask;
        }
        close_launch(L, (uint32_t)plan->tiles.size());
    }
    plan->arena_words = arena16 * 4 + 32;        // slack: kernels may read one 16-byte unit past a block
}

// 8 bases (one per byte, codes 0..4) -> 8 nibbles, first base in the low nibble
static inline uint32_t pack8(uint64_t x)
{
    x = (x | (x >> 4)) & 0x00ff00ff00ff00ffull;
    x = (x | (x >> 8)) & 0x0000ffff0000ffffull;
    x = (x | (x >> 16)) & 0x00000000ffffffffull;
    return (uint32_t)x;
}

static inline void pack_seq(const uint8_t* s, int len, int nwords, uint32_t* dst, size_t stride)
{
    const int full = len >> 3;
    for (int k = 0; k < full; ++k) {
        uint64_t x;
        memcpy(&x, s + 8 * k, 8);
        dst[(size_t)k * stride] = pack8(x);
    }
    int k = full;
    if (len & 7) {
        uint64_t x = 0;
        memcpy(&x, s + 8 * full, (size_t)(len & 7));
        dst[(size_t)k * stride] = pack8(x);
        ++k;
    }
    for (; k < nwords; ++k) dst[(size_t)k * stride] = 0;
}

void pack_arena(const ExtTask* tasks, const Plan& plan, const SchedOptions& opt, uint32_t* arena)
{
    for (const Launch& L : plan.launches) {
        const bool is_k2 = L.kind == 2;
        pfor(L.ntiles, is_k2 ? 1 : 16, opt.host_threads, [&](size_t lo, size_t hi) {
            for (size_t tix = L.tile0 + lo; tix < L.tile0 + hi; ++tix) {
                const TileHdr& hd = plan.tiles[tix];
                const int nqw = (int)(hd.nqw_ntw & 0xffffu), ntw = (int)(hd.nqw_ntw >> 16);
                uint32_t* qb = arena + (size_t)hd.qoff16 * 4;
                uint32_t* tb = arena + (size_t)hd.toff16 * 4;
                if (is_k2) {
                    const ExtTask& t = tasks[plan.slot_task[hd.slot0]];
                    pack_seq(t.q, t.qlen, (nqw + 3) & ~3, qb, 1);
                    pack_seq(t.t, t.tlen, (ntw + 3) & ~3, tb, 1);
                } else {
                    for (int lane = 0; lane < TILE_LANES; ++lane) {
                        const int64_t ti = plan.slot_task[hd.slot0 + lane];
                        if (ti < 0) {
                            for (int k = 0; k < nqw; ++k) qb[(size_t)k * TILE_LANES + lane] = 0;
                            for (int k = 0; k < ntw; ++k) tb[(size_t)k * TILE_LANES + lane] = 0;
                        } else {
                            const ExtTask& t = tasks[ti];
                            pack_seq(t.q, t.qlen, nqw, qb + lane, TILE_LANES);
                            pack_seq(t.t, t.tlen, ntw, tb + lane, TILE_LANES);
                        }
                    }
                }
            }
        });
    }
    // zero the tail slack
    const size_t used = plan.arena_words - 32;
    for (size_t k = used; k < plan.arena_words; ++k) arena[k] = 0;
}

}  // namespace bsw
