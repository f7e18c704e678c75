// bsw_chain.cpp -- host task builder: the part of BWA's mem_chain2aln either side of the extension kernel.
//
// The FPGA receives, per seed of a chain, what sw_pe_array_proc_element.v:815-934 decodes: the two query flanks and the
// two reference flanks (left ones REVERSED, pe:1638-1645 reads both forward), their lengths, h0 = seed_len * a, the
// initial score and the seed's query start.  Producing those from (read, reference window, chain) is host work in the
// quickassist port of BWA 0.7.8, which is not mounted; what follows restates the published BWA-MEM algorithm
// (mem_chain2aln / cal_max_gap of bwamem.c, 0.7.x) -- no line of it is in /root/reference.
//   bsw_chain_window      rmax[0..1]: the reference span the chain's extensions may touch
//   bsw_build_seed_tasks  one bsw_seed_task per seed (flank reversal into caller scratch, h0, init_score, qbeg)
//   bsw_finish_seed       record (relative to the seed, as the PE returns it) -> absolute query / reference coordinates
#include "../../include/bsw.h"

#include <cstring>

namespace {
// cal_max_gap: the longest gap an extension of qlen bases can afford, capped at 2w
int max_gap(const bsw_chain_opt* o, int qlen)
{
    const int l_del = (int)((double)(qlen * o->a - o->o_del) / o->e_del + 1.);
    const int l_ins = (int)((double)(qlen * o->a - o->o_ins) / o->e_ins + 1.);
    int l = l_del > l_ins ? l_del : l_ins;
    l = l > 1 ? l : 1;
    return l < (o->w << 1) ? l : (o->w << 1);
}
}  // namespace

extern "C" {

int bsw_chain_window(const bsw_chain_opt* opt, int l_query, const bsw_chain_seed* seeds, int n, int64_t l_pac, int64_t rmax[2])
{
    if (!opt || !seeds || !rmax || n < 1 || l_query < 1 || opt->e_del < 1 || opt->e_ins < 1) return BSW_EINVAL;
    rmax[0] = l_pac << 1; rmax[1] = 0;
    for (int i = 0; i < n; ++i) {
        const bsw_chain_seed& t = seeds[i];
        if (t.qbeg < 0 || t.len < 1 || t.qbeg + t.len > l_query) return BSW_EINVAL;
        const int64_t b = t.rbeg - (t.qbeg + max_gap(opt, t.qbeg));
        const int tail = l_query - t.qbeg - t.len;
        const int64_t e = t.rbeg + t.len + (tail + max_gap(opt, tail));
        rmax[0] = rmax[0] < b ? rmax[0] : b;
        rmax[1] = rmax[1] > e ? rmax[1] : e;
    }
    rmax[0] = rmax[0] > 0 ? rmax[0] : 0;
    rmax[1] = rmax[1] < (l_pac << 1) ? rmax[1] : (l_pac << 1);
    if (rmax[0] < l_pac && l_pac < rmax[1]) {           // the span crosses the forward / reverse boundary: keep the seed's side
        if (seeds[0].rbeg < l_pac) rmax[1] = l_pac;
        else rmax[0] = l_pac;
    }
    return BSW_OK;
}

size_t bsw_seed_scratch_bytes(int l_query, int64_t rmax0, int64_t rmax1, int n)
{
    return (size_t)n * ((size_t)l_query + (size_t)(rmax1 - rmax0)) + 16;
}

int bsw_build_seed_tasks(const bsw_chain_opt* opt, const uint8_t* query, int l_query, const uint8_t* rseq, int64_t rmax0,
                         int64_t rmax1, const bsw_chain_seed* seeds, int n, uint8_t* scratch, size_t scratch_bytes, bsw_seed_task* out)
{
    if (!opt || !query || !rseq || !seeds || !out || (n > 0 && !scratch) || rmax1 < rmax0) return BSW_EINVAL;
    size_t used = 0;
    for (int i = 0; i < n; ++i) {
        const bsw_chain_seed& s = seeds[i];
        bsw_seed_task& t = out[i];
        if (s.qbeg < 0 || s.len < 1 || s.qbeg + s.len > l_query || s.rbeg < rmax0 || s.rbeg + s.len > rmax1) return BSW_EINVAL;
        memset(&t, 0, sizeof(t));
        t.id = (uint32_t)i;
        t.qbeg = s.qbeg;
        t.h0 = s.len * opt->a;                                          // param word 4 (pe:826-828)
        const int64_t tl = s.rbeg - rmax0;                              // left reference flank
        const int qe = s.qbeg + s.len;
        const int64_t re = s.rbeg + s.len - rmax0, tr = (rmax1 - rmax0) - re;
        if (s.qbeg > 0) {
            // left extension: both flanks reversed (the extension runs away from the seed)
            if (tl < 1 || tl > 0x7fffffff) return BSW_EINVAL;
            if (used + (size_t)s.qbeg + (size_t)tl > scratch_bytes) return BSW_ENOMEM;
            uint8_t* qs = scratch + used; used += (size_t)s.qbeg;
            uint8_t* rs = scratch + used; used += (size_t)tl;
            for (int k = 0; k < s.qbeg; ++k) qs[k] = query[s.qbeg - 1 - k];
            for (int64_t k = 0; k < tl; ++k) rs[k] = rseq[tl - 1 - k];
            t.q_left = qs; t.t_left = rs; t.qlen[0] = s.qbeg; t.tlen[0] = (int32_t)tl;
            t.init_score = -1;                                          // a->score is unset before the left extension
        } else {
            t.init_score = s.len * opt->a;                              // no left flank: score = truesc = seed_len * a
        }
        if (qe != l_query) {
            if (tr < 1 || tr > 0x7fffffff) return BSW_EINVAL;
            t.q_right = query + qe; t.t_right = rseq + re; t.qlen[1] = l_query - qe; t.tlen[1] = (int32_t)tr;
        }
    }
    return BSW_OK;
}

// The PE's record is relative to the seed (rb/re/qe: sw_pe_array_proc_element.v:1662-1665); BWA adds the seed back.
// A seed that spans the whole read (no flank at all) comes back with score 0 / truesc = init_score (pe:717-719,757-759):
// BWA sets score = truesc = seed_len * a there.
void bsw_finish_seed(const bsw_chain_opt* opt, const bsw_chain_seed* s, int l_query, const bsw_aln_record* r, bsw_seed_aln* a)
{
    const int qe_seed = s->qbeg + s->len;
    a->score = r->score; a->truesc = r->truesc; a->w = r->w;
    if (s->qbeg > 0) { a->qb = r->qb; a->rb = s->rbeg + r->rb; }       // rb <= 0: reference bases left of the seed
    else { a->qb = 0; a->rb = s->rbeg; }
    if (qe_seed != l_query) { a->qe = qe_seed + r->qe; a->re = s->rbeg + s->len + r->re; }
    else { a->qe = l_query; a->re = s->rbeg + s->len; }
    if (s->qbeg == 0 && qe_seed == l_query) a->score = a->truesc = s->len * opt->a;
    a->seedcov = s->len;
}

}  // extern "C"
