// synth.cpp -- deterministic synthetic seed-extension task generator (workload definition for
// BASELINE.json configs 1-5).  Built as libbsw_synth.so; used by bench.py and
// tests to create inputs.  No reference code involved: the reference tree has no generator.
//
// Model: a read of length L carries one exact seed (seed_len ~ U[seed_min, seed_max], placed so that
// both flanks are non-empty).  Read r yields task 2r (left flank, reversed, h0 = seed_len*a) and
// task 2r+1 (right flank, h0 = (seed_len + left_len)*a, i.e. the score of a clean left extension).
// The target is the flank with substitutions / insertions / deletions applied, padded with random
// bases to tlen = qlen + min(max(1, qlen*a - o + 1), 2w)  (BWA's cal_max_gap window), optionally
// switched to unrelated sequence after a random breakpoint.  "Long" mode draws qlen directly.
// Every task is generated from its own counter-based RNG stream (seed, task index), so any chunk
// [first, first+n) can be regenerated independently (config 5 streams 100M tasks this way).
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include <algorithm>

extern "C" {

typedef struct {
    int32_t read_len_min, read_len_max;   // short-read mode: L ~ U[min,max]
    int32_t seed_min, seed_max;           // seed_len ~ U[seed_min, min(seed_max, L-2)]
    int32_t long_mode;                    // 1: qlen ~ U[qlen_min,qlen_max], h0 ~ U[h0_min,h0_max]
    int32_t qlen_min, qlen_max, h0_min, h0_max;
    int32_t w, a, o;                      // band, match score, gap open (for the target window)
    double  sub, ins, del;                // per-base rates; ins = base present in query only
    double  unrelated_frac;               // fraction of tasks whose target turns random after a breakpoint
    double  n_frac;                       // per-base probability of an ambiguous base (code 4) in the query
    uint64_t seed;
} bsw_synth_cfg;

}  // extern "C"

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed, uint64_t stream) { s = seed * 0x9E3779B97F4A7C15ull + stream * 0xD1B54A32D192ED03ull + 0x8CB92BA72F3D8DD7ull; next(); next(); }
    inline uint64_t next() {  // splitmix64
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    inline uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }   // [0,n)
    inline int range(int lo, int hi) { return hi <= lo ? lo : lo + (int)below((uint32_t)(hi - lo + 1)); }  // [lo,hi]
    inline double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

struct Shape { int qlen, tlen, h0; };

// Lengths and h0 depend only on the first draws of the task-pair stream.
inline Shape task_shape(const bsw_synth_cfg& c, int64_t task)
{
    Shape sh;
    if (c.long_mode) {
        Rng r(c.seed, (uint64_t)task * 2 + 1);
        sh.qlen = r.range(c.qlen_min, c.qlen_max);
        sh.h0 = r.range(c.h0_min, c.h0_max);
    } else {
        Rng r(c.seed, (uint64_t)(task >> 1) * 2);          // per-read stream
        int L = r.range(c.read_len_min, c.read_len_max);
        if (L < 3) L = 3;
        int smax = std::min(c.seed_max, L - 2), smin = std::min(c.seed_min, smax);
        int seed_len = r.range(smin, smax);
        int start = r.range(1, L - seed_len - 1);          // both flanks non-empty
        int left = start, right = L - start - seed_len;
        if ((task & 1) == 0) { sh.qlen = left;  sh.h0 = seed_len * c.a; }
        else                 { sh.qlen = right; sh.h0 = (seed_len + left) * c.a; }
    }
    int gap = sh.qlen * c.a - c.o + 1;
    if (gap < 1) gap = 1;
    if (gap > 2 * c.w) gap = 2 * c.w;
    sh.tlen = sh.qlen + gap;
    return sh;
}

inline void task_fill(const bsw_synth_cfg& c, int64_t task, const Shape& sh, uint8_t* q, uint8_t* t)
{
    Rng r(c.seed ^ 0xA5A5A5A55A5A5A5Aull, (uint64_t)task);
    for (int j = 0; j < sh.qlen; ++j) q[j] = (uint8_t)(r.next() >> 62);
    int brk = sh.tlen + 1;
    if (c.unrelated_frac > 0 && r.unit() < c.unrelated_frac) brk = r.range(0, sh.tlen - 1);
    int ti = 0, j = 0;
    while (ti < sh.tlen) {
        if (ti >= brk || j >= sh.qlen) { t[ti++] = (uint8_t)(r.next() >> 62); continue; }
        double u = r.unit();
        if (u < c.sub) { t[ti++] = (uint8_t)((q[j] + 1 + r.below(3)) & 3); ++j; }
        else if (u < c.sub + c.ins) { ++j; }                                   // base only in the query
        else if (u < c.sub + c.ins + c.del) {                                  // extra target bases, geometric p=0.5
            do { t[ti++] = (uint8_t)(r.next() >> 62); } while (ti < sh.tlen && (r.next() >> 63));
        } else { t[ti++] = q[j]; ++j; }
    }
    if (c.n_frac > 0)
        for (int k = 0; k < sh.qlen; ++k) if (r.unit() < c.n_frac) q[k] = 4;
}

template <class F>
void parallel_for(int64_t n, F f)
{
    unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    if (n < 4096) nt = 1;
    nt = std::min<unsigned>(nt, 64);
    std::vector<std::thread> th;
    int64_t per = (n + nt - 1) / nt;
    for (unsigned k = 0; k < nt; ++k) {
        int64_t lo = k * per, hi = std::min<int64_t>(n, lo + per);
        if (lo >= hi) break;
        th.emplace_back([=] { f(lo, hi); });
    }
    for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

// Pass 1: shapes of tasks [first, first+n).  qlen/tlen/h0 arrays of n int32.
void bsw_synth_shapes(const bsw_synth_cfg* cfg, int64_t first, int64_t n, int32_t* qlen, int32_t* tlen, int32_t* h0)
{
    bsw_synth_cfg c = *cfg;
    parallel_for(n, [=](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            Shape s = task_shape(c, first + i);
            qlen[i] = s.qlen; tlen[i] = s.tlen; h0[i] = s.h0;
        }
    });
}

// Pass 2: bases.  qoff/toff are the exclusive prefix sums of qlen/tlen (n+1 entries each).
void bsw_synth_fill(const bsw_synth_cfg* cfg, int64_t first, int64_t n, const int64_t* qoff, const int64_t* toff,
                    uint8_t* qbuf, uint8_t* tbuf)
{
    bsw_synth_cfg c = *cfg;
    parallel_for(n, [=](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            Shape s = task_shape(c, first + i);
            task_fill(c, first + i, s, qbuf + qoff[i], tbuf + toff[i]);
        }
    });
}

}  // extern "C"
