// bsw_pack2.cpp -- 2 bit per base packer of the lean flat path (north_star: "2-bit/4-bit packed sequences").
// One streaming pass over a chunk of a flat batch: every sequence is packed 16 bases per u32 (base j in bits
// 2*(j & 15) of word j >> 4), sequences start on a word; a task that holds an N (code 4) cannot be expressed and is
// reported instead (the caller reruns it on the 4-bit path), a code above 4 ends the pass with the task's index.
// Host->device bytes per 150 bp task: ~52 (+24 of scalars) against 190 when the DMA engine copies one byte per base --
// with eight GPUs on one box the copies, not the kernels, bound the batch call.
// Compiled twice like bsw_pack.cpp: BSW_PACK_ISA=0 (portable) and 512 (-mavx512bw -mavx512vl -mbmi2).
#include "bsw_sched.h"

#include <cstring>
#include <immintrin.h>

namespace bsw {

#if BSW_PACK_ISA == 512
#define PACK2_NAME pack2_flat_avx512
// 64 bases per step: b0 + 4 b1 per 16-bit lane (vpmaddubsw), x0 + 16 x1 per 32-bit lane (vpmaddwd), low bytes (vpmovdb).
// The masked load reads exactly the sequence; the 16-byte store may run past it into the next sequence's words, which
// are written afterwards (the arena is filled front to back and carries 64 bytes of slack).
static inline void pack2_step(const uint8_t* s, int rem, uint8_t* d, __m512i* mx)
{
    rem = rem < 0 ? 0 : (rem > 64 ? 64 : rem);
    const __m512i x = _mm512_maskz_loadu_epi8(_bzhi_u64(~0ull, (unsigned)rem), s);
    *mx = _mm512_max_epu8(*mx, x);
    const __m512i a = _mm512_maddubs_epi16(x, _mm512_set1_epi16(0x0401));
    const __m512i b = _mm512_madd_epi16(a, _mm512_set1_epi32(0x00100001));
    _mm_storeu_si128(reinterpret_cast<__m128i*>(d), _mm512_cvtepi32_epi8(b));
}
struct Pack2Acc {
    __m512i m;
    Pack2Acc() : m(_mm512_setzero_si512()) {}
    void seq(const uint8_t* s, int len, uint32_t* dst)
    {
        uint8_t* d = reinterpret_cast<uint8_t*>(dst);
        if (len <= 128) { pack2_step(s, len, d, &m); pack2_step(s + 64, len - 64, d + 16, &m); return; }
        for (int done = 0; done < len; done += 64) pack2_step(s + done, len - done, d + (done >> 2), &m);
    }
    int worst()                                   // 0: all codes <= 3, else the largest code seen
    {
        if (!_mm512_cmpgt_epu8_mask(m, _mm512_set1_epi8(3))) return 0;
        alignas(64) uint8_t b[64];
        _mm512_store_si512(reinterpret_cast<__m512i*>(b), m);
        int w = 0;
        for (int k = 0; k < 64; ++k) w = b[k] > w ? b[k] : w;
        return w;
    }
    void reset() { m = _mm512_setzero_si512(); }
};
#else
#define PACK2_NAME pack2_flat_generic
struct Pack2Acc {
    int mx = 0;
    void seq(const uint8_t* s, int len, uint32_t* dst)
    {
        const int nw = (len + 15) >> 4;
        for (int wd = 0; wd < nw; ++wd) {
            uint32_t v = 0;
            const int lim = len - 16 * wd < 16 ? len - 16 * wd : 16;
            for (int k = 0; k < lim; ++k) { const int c = s[16 * wd + k]; mx = c > mx ? c : mx; v |= (uint32_t)(c & 3) << (2 * k); }
            dst[wd] = v;
        }
    }
    int worst() { return mx > 3 ? mx : 0; }
    void reset() { mx = 0; }
};
#endif

// tasks [0, count): query qbuf[qoff[i] .. qoff[i+1]), target likewise.  src[i] = word offsets of the two packed
// sequences in `arena`; n_list receives the indices of tasks that hold an N.  Returns the words used, or -1 - i when
// task i holds a code above 4.
int64_t PACK2_NAME(const uint8_t* qbuf, const int64_t* qoff, const uint8_t* tbuf, const int64_t* toff, size_t count,
                   uint32_t* arena, SlotSrc* src, std::vector<uint32_t>* n_list)
{
    size_t w = 0;
    Pack2Acc acc;
    for (size_t i = 0; i < count; ++i) {
        const int ql = (int)(qoff[i + 1] - qoff[i]), tl = (int)(toff[i + 1] - toff[i]);
        if (i + 4 < count) {
            _mm_prefetch(reinterpret_cast<const char*>(qbuf + qoff[i + 4]), _MM_HINT_T0);
            _mm_prefetch(reinterpret_cast<const char*>(tbuf + toff[i + 4]), _MM_HINT_T0);
            _mm_prefetch(reinterpret_cast<const char*>(tbuf + toff[i + 4] + 64), _MM_HINT_T0);
        }
        src[i].qoff16 = (uint32_t)w;
        acc.seq(qbuf + qoff[i], ql, arena + w);
        w += (size_t)(ql + 15) >> 4;
        src[i].toff16 = (uint32_t)w;
        acc.seq(tbuf + toff[i], tl, arena + w);
        w += (size_t)(tl + 15) >> 4;
        const int worst = acc.worst();
        if (worst) {
            if (worst > 4) return -1 - (int64_t)i;
            n_list->push_back((uint32_t)i);
            acc.reset();
        }
    }
    return (int64_t)w;
}

}  // namespace bsw
