// bsw_pack.cpp -- the streaming pass over the caller's bases: validate the codes, detect N, pack 4 bit/base into the
// task-major source arena (sw_pe_array_query_mem.v / task_parse's 4-bit base stream).
// Compiled twice (see the Makefile): BSW_PACK_ISA=0 -> pack_tasks_sse2 (x86-64 baseline), BSW_PACK_ISA=512 ->
// pack_tasks_avx512 (-mavx512bw -mavx512vl -mbmi2); bsw_sched.cpp picks one at run time.
#include "bsw_sched.h"
#include "../../include/bsw.h"

#include <algorithm>
#include <cstring>
#include <immintrin.h>

namespace bsw {

#ifndef PACK_FIXED
#define PACK_FIXED 2     // measured: fixed two steps up to 128 bases, four up to 256 (34 ns/task vs 42 with a counted loop)
#endif
#ifndef PACK_PREFETCH
#define PACK_PREFETCH 8
#endif

#if BSW_PACK_ISA == 512
#define PACK_TASKS_NAME pack_tasks_avx512
typedef __m512i MaxAcc;
static inline MaxAcc max_zero() { return _mm512_setzero_si512(); }
static inline int max_reduce(MaxAcc v)
{
    const __m256i a = _mm256_max_epu8(_mm512_castsi512_si256(v), _mm512_extracti64x4_epi64(v, 1));
    __m128i m = _mm_max_epu8(_mm256_castsi256_si128(a), _mm256_extracti128_si256(a, 1));
    m = _mm_max_epu8(m, _mm_srli_si128(m, 8)); m = _mm_max_epu8(m, _mm_srli_si128(m, 4));
    m = _mm_max_epu8(m, _mm_srli_si128(m, 2)); m = _mm_max_epu8(m, _mm_srli_si128(m, 1));
    return _mm_cvtsi128_si32(m) & 0xff;
}

// 64 bases per step.  Masked loads read exactly the sequence (masked-off bytes cannot fault), masked stores write
// ceil(rem/32)*16 bytes, so a step needs no tail or page-boundary branch; sequences up to 128 / 256 bases (the bulk
// of short-read extension) take a fixed two / four steps, which removes the data-dependent loop exit as well.
static inline void pack_step(const uint8_t* s, int len, int done, uint8_t* dst, __m512i* m)
{
    int rem = len - done;
    rem = rem < 0 ? 0 : (rem > 64 ? 64 : rem);
    const __mmask64 lm = _bzhi_u64(~0ull, (unsigned)rem);
    const __m512i x = _mm512_maskz_loadu_epi8(lm, s + done);
    *m = _mm512_max_epu8(*m, x);
    // per 16-bit lane {b1,b0}: low byte of x | x>>4 = b0 | b1<<4; vpmovwb keeps exactly that byte
    const __m256i z = _mm512_cvtepi16_epi8(_mm512_or_si512(x, _mm512_srli_epi16(x, 4)));
    const __mmask32 sm = rem > 32 ? 0xffffffffu : (rem > 0 ? 0xffffu : 0u);
    _mm256_mask_storeu_epi8(dst + (done >> 1), sm, z);
}
static inline int pack_seq(const uint8_t* s, int len, uint32_t* dst, MaxAcc* mx)
{
    uint8_t* d = reinterpret_cast<uint8_t*>(dst);
    __m512i m = *mx;
#if PACK_FIXED == 4
    if (len <= 256) {
        pack_step(s, len, 0, d, &m); pack_step(s, len, 64, d, &m); pack_step(s, len, 128, d, &m); pack_step(s, len, 192, d, &m);
    } else
#elif PACK_FIXED == 2
    if (len <= 128) {
        pack_step(s, len, 0, d, &m); pack_step(s, len, 64, d, &m);
    } else if (len <= 256) {
        pack_step(s, len, 0, d, &m); pack_step(s, len, 64, d, &m); pack_step(s, len, 128, d, &m); pack_step(s, len, 192, d, &m);
    } else
#endif
    {
        for (int done = 0; done < len; done += 64) pack_step(s, len, done, d, &m);
    }
    *mx = m;
    return ((len + 31) >> 5) << 2;
}
#else
#define PACK_TASKS_NAME pack_tasks_sse2
typedef __m128i MaxAcc;
static inline MaxAcc max_zero() { return _mm_setzero_si128(); }
static inline int max_reduce(MaxAcc m)
{
    m = _mm_max_epu8(m, _mm_srli_si128(m, 8)); m = _mm_max_epu8(m, _mm_srli_si128(m, 4));
    m = _mm_max_epu8(m, _mm_srli_si128(m, 2)); m = _mm_max_epu8(m, _mm_srli_si128(m, 1));
    return _mm_cvtsi128_si32(m) & 0xff;
}
// Packs `len` bases into dst (ceil(len/8) words, then zero-padded to a multiple of 4 words).  Returns the number of
// words written (multiple of 4); *mx accumulates the bytewise maximum of the codes (validation: max <= 4, and max == 4
// <=> the task holds an N), so the check costs one op per 16 bases.
alignas(16) static const uint8_t k_tail_mask[32] = { 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255,
                                                     0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
static inline __m128i nib16(__m128i x, __m128i lowbyte)
{
    // per 16-bit lane {b1,b0} -> b0 | b1<<4 in the low byte
    return _mm_and_si128(_mm_or_si128(x, _mm_srli_epi16(x, 4)), lowbyte);
}
static inline int pack_seq(const uint8_t* s, int len, uint32_t* dst, MaxAcc* mx)
{
    const __m128i lowbyte = _mm_set1_epi16(0x00ff);
    __m128i m = *mx;
    int k = 0, done = 0;
    for (; done + 32 <= len; done += 32, k += 4) {
        const __m128i x0 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + done));
        const __m128i x1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + done + 16));
        m = _mm_max_epu8(m, _mm_max_epu8(x0, x1));
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + k), _mm_packus_epi16(nib16(x0, lowbyte), nib16(x1, lowbyte)));
    }
    // the last 1..31 bases: two 16-byte loads masked to the sequence.  Reading past the end is harmless as long as the
    // load stays inside the page of a byte we own; next to a page end the bytes go through a bounce buffer instead.
    const int rem = len - done;
    if (rem > 0) {
        __m128i x0, x1;
        if (((uintptr_t)(s + done) & 4095u) <= 4096u - 32u) {
            x0 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + done));
            x1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + done + 16));
        } else {
            alignas(16) uint8_t buf[32];
            for (int j = 0; j < 32; ++j) buf[j] = j < rem ? s[done + j] : 0;
            x0 = _mm_load_si128(reinterpret_cast<const __m128i*>(buf));
            x1 = _mm_load_si128(reinterpret_cast<const __m128i*>(buf + 16));
        }
        const int r0 = rem < 16 ? rem : 16, r1 = rem - r0;
        x0 = _mm_and_si128(x0, _mm_loadu_si128(reinterpret_cast<const __m128i*>(k_tail_mask + 16 - r0)));
        x1 = _mm_and_si128(x1, _mm_loadu_si128(reinterpret_cast<const __m128i*>(k_tail_mask + 16 - r1)));
        m = _mm_max_epu8(m, _mm_max_epu8(x0, x1));
        _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + k), _mm_packus_epi16(nib16(x0, lowbyte), nib16(x1, lowbyte)));
        k += 4;
    }
    *mx = m;
    return k;
}

#endif

int PACK_TASKS_NAME(const ExtTask* tasks, size_t n, int max_mat, const SchedOptions& opt, uint8_t* cls, SlotSrc* src,
               uint32_t* arena, size_t* words_used, size_t* bad_task, std::string* msg)
{
    const int k1cap = K1_QLEN_CAP;
    size_t w = 0;                                       // next free word (multiple of 4)
    for (size_t i = 0; i < n; ++i) {
        const ExtTask& t = tasks[i];
        if (i + PACK_PREFETCH < n) {                       // the bases stream in from DRAM: pull a few tasks ahead
            const ExtTask& a = tasks[i + PACK_PREFETCH];
            _mm_prefetch(reinterpret_cast<const char*>(a.q), _MM_HINT_T0); _mm_prefetch(reinterpret_cast<const char*>(a.q) + 64, _MM_HINT_T0);
            _mm_prefetch(reinterpret_cast<const char*>(a.t), _MM_HINT_T0); _mm_prefetch(reinterpret_cast<const char*>(a.t) + 64, _MM_HINT_T0);
            _mm_prefetch(reinterpret_cast<const char*>(a.t) + 128, _MM_HINT_T0);
        }
        if (t.qlen == 0 && t.w == -2) { cls[i] = 0x80; src[i] = SlotSrc{ 0, 0 }; continue; }     // absent flank of a seed task
        int e = 0;
        if (!t.q || !t.t || t.qlen < 1 || t.tlen < 1 || t.h0 < 1 || t.w < 0) e = BSW_EINVAL;
        else if ((int64_t)t.h0 + (int64_t)t.qlen * max_mat > SCORE_CAP || t.qlen > K2_QLEN_CAP || t.tlen > 500000) e = BSW_ERANGE;
        uint8_t c = 0;
        if (!e) {
            MaxAcc mx = max_zero();
            src[i].qoff16 = (uint32_t)(w >> 2);
            w += (size_t)pack_seq(t.q, t.qlen, arena + w, &mx);
            src[i].toff16 = (uint32_t)(w >> 2);
            w += (size_t)pack_seq(t.t, t.tlen, arena + w, &mx);
            const int top = max_reduce(mx);                  // largest base code of the task
            if (top > 4) e = BSW_EINVAL;
            if (top == 4 || !opt.fast_matrix) c |= 1;        // an N: matrix-lookup scoring
            bool longtask = opt.force_kernel == 2 || (opt.force_kernel == 0 && t.qlen >= opt.k2_min_qlen) || t.qlen > k1cap;
            if (longtask) c |= 2;
        }
        cls[i] = c;
        if (e) {
            if (bad_task) *bad_task = i;
            if (msg)
                *msg = "task " + std::to_string(i) + ": qlen=" + std::to_string(t.qlen) + " tlen=" + std::to_string(t.tlen) +
                       " h0=" + std::to_string(t.h0) + " w=" + std::to_string(t.w) +
                       (e == BSW_ERANGE ? " outside the 16-bit envelope (h0 + qlen * max(mat) <= 32767, qlen <= 40000, tlen <= 500000)"
                                        : " invalid (null pointer, length < 1, h0 < 1 or base code > 4)");
            return e;
        }
    }
    for (int k = 0; k < 8; ++k) arena[w + (size_t)k] = 0;     // slack: the gather may read one 16-byte unit past a sequence
    *words_used = w + 8;
    return 0;
}


}  // namespace bsw
