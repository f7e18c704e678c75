// bsw_wire.cpp -- FPGA wire formats: task batch buffer (TBB) in, result batch buffer (RBB) out.
//
// Layouts derived from the RTL:
//   TBB  65 536 u32 = 4096 x 64 B (bwa_mem_sw.v:163-166, tbb.v:163-194): header words 0-2 (proc_element.v:815-820,
//        915-918; task_parse.v:944), 8 parameter words per task at 8+8i (proc_element.v:807,826-828,871-874,880-892,
//        924-934), then the packed bases, 8 per word, first base in bits [31:28] (proc_element.v:1638,1677), segments
//        qL(rev) qR tL(rev) tR back to back (task_parse.v:1900,1896), task i's data at 8+8N+(off_i-off_0)
//        (task_parse.v:1928-1929,1924,1936).
//   RBB  4 096 u32 = 256 x 64 B (bwa_mem_sw.v:167-170, rbb.v:117-167): 5 words per task
//        [id][qe<<16|qb][re<<16|rb][truesc<<16|score][w] (proc_element.v:1187-1205,1662-1665; fill_resulBuf.v:377-429).
// The FPGA emits records in completion order; bsw_fpga_batch emits them in task order, which is one valid completion
// order (records are identified by word 0).
#include <cstring>
#include <string>
#include <vector>

#include "bsw_internal.h"
#include "bsw_sched.h"

namespace {

inline uint32_t get_base(const uint32_t* data, size_t k)      // k-th base of a task's data block, MS nibble first
{
    return (data[k >> 3] >> (28 - 4 * (k & 7))) & 15u;
}

}  // namespace

extern "C" {

int bsw_tbb_encode(const bsw_params2* P, const bsw_seed_task* tasks, size_t n, uint32_t* tbb)
{
    if (!P || !tbb || (!tasks && n)) return BSW_EINVAL;
    if (n > (size_t)(BSW_RBB_WORDS / 5)) return BSW_EWIRE;                           // fill_resulBuf.v:378: <= 819 records
    const bsw_params& p = P->p;
    if (p.o_del < 0 || p.o_del > 255 || p.e_del < 0 || p.e_del > 255 || p.o_ins < 0 || p.o_ins > 255 || p.e_ins < 0 ||
        p.e_ins > 255 || P->pen_clip5 < 0 || P->pen_clip5 > 255 || P->pen_clip3 < 0 || P->pen_clip3 > 255 || P->w < 0 || P->w > 255)
        return BSW_EWIRE;
    memset(tbb, 0, sizeof(uint32_t) * BSW_TBB_WORDS);
    tbb[0] = (uint32_t)p.o_del | ((uint32_t)p.e_del << 8) | ((uint32_t)p.o_ins << 16) | ((uint32_t)p.e_ins << 24);
    tbb[1] = (uint32_t)P->pen_clip5 | ((uint32_t)P->pen_clip3 << 8) | ((uint32_t)P->w << 16);
    tbb[2] = (uint32_t)n;
    size_t off = 8 + 8 * n;                                                          // word offset of the next data block
    for (size_t i = 0; i < n; ++i) {
        const bsw_seed_task& s = tasks[i];
        if (s.qlen[0] < 0 || s.qlen[0] > 255 || s.qlen[1] < 0 || s.qlen[1] > 255 || s.tlen[0] < 0 || s.tlen[0] > 2047 ||
            s.tlen[1] < 0 || s.tlen[1] > 2047 || s.h0 < 0 || s.h0 > 255 || s.init_score < -32768 || s.init_score > 32767 ||
            s.qbeg < 0 || s.qbeg > 65535)
            return BSW_EWIRE;
        const size_t nb = (size_t)s.qlen[0] + s.qlen[1] + s.tlen[0] + s.tlen[1];
        if (nb > 2048) return BSW_EWIRE;                                             // query_mem 2048 x 4 b (proc_element.v:347-350)
        const size_t nw = (nb + 7) / 8;
        if (off + nw > BSW_TBB_WORDS) return BSW_EWIRE;
        uint32_t* pw = tbb + 8 + 8 * i;
        pw[0] = (uint32_t)s.qlen[0] | ((uint32_t)s.tlen[0] << 16);
        pw[1] = (uint32_t)s.qlen[1] | ((uint32_t)s.tlen[1] << 16);
        pw[2] = (uint32_t)off;
        pw[3] = ((uint32_t)s.init_score & 0xffffu) | ((uint32_t)s.qbeg << 16);
        pw[4] = (uint32_t)s.h0;
        for (int side = 0; side < 2; ++side) {
            // the host precomputes ksw_extend2's max_ins / max_del (end_bonus = pen_clip5 / pen_clip3)
            const int eb = side ? P->pen_clip3 : P->pen_clip5;
            int mx = 0;
            for (int k = 0; k < 25; ++k) mx = mx > p.mat[k] ? mx : p.mat[k];
            int max_ins = 1, max_del = 1;
            if (p.e_ins > 0) { max_ins = (int)((double)(s.qlen[side] * mx + eb - p.o_ins) / p.e_ins + 1.); if (max_ins < 1) max_ins = 1; }
            if (p.e_del > 0) { max_del = (int)((double)(s.qlen[side] * mx + eb - p.o_del) / p.e_del + 1.); if (max_del < 1) max_del = 1; }
            if (max_ins > 65535) max_ins = 65535;
            if (max_del > 65535) max_del = 65535;
            pw[5 + side] = (uint32_t)max_ins | ((uint32_t)max_del << 16);
        }
        pw[7] = s.id;
        const uint8_t* seg[4] = { s.q_left, s.q_right, s.t_left, s.t_right };
        const int len[4] = { s.qlen[0], s.qlen[1], s.tlen[0], s.tlen[1] };
        size_t k = 0;
        for (int g = 0; g < 4; ++g) {
            if (len[g] && !seg[g]) return BSW_EINVAL;
            for (int j = 0; j < len[g]; ++j, ++k) {
                if (seg[g][j] > 4) return BSW_EINVAL;
                tbb[off + (k >> 3)] |= (uint32_t)seg[g][j] << (28 - 4 * (k & 7));
            }
        }
        off += nw;
    }
    return BSW_OK;
}

// The envelope in which the FPGA's 8-bit datapath is exact (oracle.rtl_envelope, checked against the
// translated RTL in tests/test_rtl_pin.py and tests/test_rtl_width_model.py): per extension
//   scores  h0 + qlen <= 127,  columns qlen <= 127,  first column h0 - o_del - e_del*tlen >= -128,
//   first row h0 - o_ins - e_ins*qlen >= -128,  band 0 < w <= 63 (w << 1 is a signed 8-bit value in the second try).
// The right extension starts from the left score, which is at most max(h0, init_score) + qlen_left.
// *n_outside = tasks for which the FPGA may return something else than ksw_extend2 (this library returns ksw_extend2's
// answer for them); *first_outside = index of the first one or -1.
int bsw_fpga_envelope(const uint32_t* tbb, int* n_outside, int* first_outside)
{
    if (!tbb || !n_outside) return BSW_EINVAL;
    *n_outside = 0;
    if (first_outside) *first_outside = -1;
    const size_t n = tbb[2];
    if (n > (size_t)(BSW_RBB_WORDS / 5) || 8 + 8 * n > BSW_TBB_WORDS) return BSW_EWIRE;
    const int o_del = (int)(tbb[0] & 0xff), e_del = (int)((tbb[0] >> 8) & 0xff), o_ins = (int)((tbb[0] >> 16) & 0xff), e_ins = (int)(tbb[0] >> 24);
    const int w = (int)((tbb[1] >> 16) & 0xff);
    for (size_t i = 0; i < n; ++i) {
        const uint32_t* pw = tbb + 8 + 8 * i;
        const int ql[2] = { (int)(pw[0] & 0xff), (int)(pw[1] & 0xff) }, tl[2] = { (int)((pw[0] >> 16) & 0x7ff), (int)((pw[1] >> 16) & 0x7ff) };
        const int h0 = (int)(pw[4] & 0xff), init = (int)(int16_t)(pw[3] & 0xffff);
        bool ok = w > 0 && w <= 63;
        int h = h0;
        for (int side = 0; side < 2 && ok; ++side) {
            if (side == 1) h = ql[0] ? h0 + ql[0] : (init > h0 ? init : h0);          // upper bound of the left score
            if (ql[side] == 0) continue;
            ok = ql[side] <= 127 && h + ql[side] <= 127 && h > 0 && h - o_del - e_del * tl[side] >= -128 &&
                 h - o_ins - e_ins * ql[side] >= -128;
        }
        if (!ok) { if (first_outside && *n_outside == 0) *first_outside = (int)i; ++*n_outside; }
    }
    return BSW_OK;
}

int bsw_rbb_decode(const uint32_t* rbb, size_t n, bsw_aln_record* out)
{
    if (!rbb || (!out && n)) return BSW_EINVAL;
    if (n > (size_t)(BSW_RBB_WORDS / 5)) return BSW_EWIRE;
    for (size_t k = 0; k < n; ++k) {
        const uint32_t* r = rbb + 5 * k;
        bsw_aln_record& o = out[k];
        o.id = r[0];
        o.qb = (int16_t)(r[1] & 0xffffu); o.qe = (int16_t)(r[1] >> 16);
        o.rb = (int16_t)(r[2] & 0xffffu); o.re = (int16_t)(r[2] >> 16);
        o.score = (int16_t)(r[3] & 0xffffu); o.truesc = (int16_t)(r[3] >> 16);
        o.w = (int32_t)r[4];
    }
    return BSW_OK;
}

int bsw_fpga_batch(bsw_ctx* ctx, const uint32_t* tbb, uint32_t* rbb, int* n_results)
{
    if (!ctx || !tbb || !rbb) return BSW_EINVAL;
    if (n_results) *n_results = 0;
    const size_t n = tbb[2];                                                         // task_parse.v:944,697
    if (n > (size_t)(BSW_RBB_WORDS / 5) || 8 + 8 * n > BSW_TBB_WORDS) { bsw_set_error_text(ctx, "TBB: task count out of range"); return BSW_EWIRE; }
    if (n == 0) return BSW_OK;
    if (bsw_option_value(ctx, "fpga_strict") > 0) {                                  // refuse what the FPGA itself cannot compute exactly
        int outside = 0, first = -1;
        const int rc = bsw_fpga_envelope(tbb, &outside, &first);
        if (rc) return rc;
        if (outside) {
            bsw_set_error_text(ctx, ("TBB: " + std::to_string(outside) + " task(s) outside the FPGA's 8-bit envelope, first at index " + std::to_string(first)).c_str());
            return BSW_ERANGE;
        }
    }
    bsw_params2 P;
    memset(&P, 0, sizeof(P));
    // the RTL's matrix is hard-wired: match +1, mismatch -4, N -1 (sw_pe_array_sw_extend.v:1915-1940)
    for (int t = 0; t < 5; ++t)
        for (int q = 0; q < 5; ++q) P.p.mat[5 * t + q] = (int8_t)((t == 4 || q == 4) ? -1 : (t == q ? 1 : -4));
    P.p.o_del = (int)(tbb[0] & 0xff); P.p.e_del = (int)((tbb[0] >> 8) & 0xff);       // proc_element.v:815-820
    P.p.o_ins = (int)((tbb[0] >> 16) & 0xff); P.p.e_ins = (int)(tbb[0] >> 24);
    P.pen_clip5 = (int)(tbb[1] & 0xff); P.pen_clip3 = (int)((tbb[1] >> 8) & 0xff);   // proc_element.v:915-918
    P.w = (int)((tbb[1] >> 16) & 0xff);
    P.p.zdrop = 0;                                                                   // the RTL has no z-drop (ports sw_extend.v:96-116)
    P.p.end_bonus = 0;                                                               // unused: max_ins/max_del come from the batch
    if (P.p.e_del < 1 || P.p.e_ins < 1) { bsw_set_error_text(ctx, "TBB: gap extension penalty is zero"); return BSW_EWIRE; }
    std::vector<bsw_seed_task> tasks(n);
    std::vector<bsw_seed_clamp> clamps(n);
    std::vector<uint8_t> bases;
    std::vector<size_t> boff(n + 1, 0);
    const size_t off0 = tbb[8 + 2];                                                  // task_parse.v:1928-1929
    for (size_t i = 0; i < n; ++i) {
        const uint32_t* pw = tbb + 8 + 8 * i;
        boff[i + 1] = boff[i] + (pw[0] & 0xff) + ((pw[0] >> 16) & 0x7ff) + (pw[1] & 0xff) + ((pw[1] >> 16) & 0x7ff);
    }
    bases.resize(boff[n] + 8);
    for (size_t i = 0; i < n; ++i) {
        const uint32_t* pw = tbb + 8 + 8 * i;
        bsw_seed_task& s = tasks[i];
        s.qlen[0] = (int)(pw[0] & 0xff); s.tlen[0] = (int)((pw[0] >> 16) & 0x7ff);   // proc_element.v:880-883
        s.qlen[1] = (int)(pw[1] & 0xff); s.tlen[1] = (int)((pw[1] >> 16) & 0x7ff);   // proc_element.v:889-892
        s.init_score = (int16_t)(pw[3] & 0xffff); s.qbeg = (int)(pw[3] >> 16);   // regScore is BWA's a->score before the task (-1 = unset)           // proc_element.v:871-874
        s.h0 = (int)(pw[4] & 0xff);                                                  // proc_element.v:826-828
        clamps[i].max_ins[0] = (int)(pw[5] & 0xffff); clamps[i].max_del[0] = (int)(pw[5] >> 16);   // proc_element.v:924-926
        clamps[i].max_ins[1] = (int)(pw[6] & 0xffff); clamps[i].max_del[1] = (int)(pw[6] >> 16);   // proc_element.v:932-934
        s.id = pw[7];                                                                // proc_element.v:807
        const size_t nb = boff[i + 1] - boff[i];
        if (nb > 2048) { bsw_set_error_text(ctx, "TBB: a task holds more than 2048 bases (query_mem, proc_element.v:347-350)"); return BSW_EWIRE; }
        const size_t dpos = 8 + 8 * n + ((size_t)pw[2] - off0);                      // task_parse.v:1924,1936
        if (pw[2] < off0 || dpos + (nb + 7) / 8 > BSW_TBB_WORDS) { bsw_set_error_text(ctx, "TBB: data offset out of range"); return BSW_EWIRE; }
        uint8_t* b = bases.data() + boff[i];
        for (size_t k = 0; k < nb; ++k) {
            const uint32_t v = get_base(tbb + dpos, k) & 7u;                         // low 3 bits used (sx:1883,1885)
            if (v > 4) { bsw_set_error_text(ctx, "TBB: base code > 4"); return BSW_EWIRE; }
            b[k] = (uint8_t)v;
        }
        s.q_left = b; s.q_right = b + s.qlen[0]; s.t_left = s.q_right + s.qlen[1]; s.t_right = s.t_left + s.tlen[0];
        if ((s.qlen[0] && (s.tlen[0] < 1 || s.h0 < 1)) || (s.qlen[1] && s.tlen[1] < 1)) {
            bsw_set_error_text(ctx, "TBB: extension with empty target or h0 == 0"); return BSW_EWIRE;
        }
    }
    std::vector<bsw_aln_record> rec(n);
    const int rc = bsw_chain2aln_impl(ctx, &P, tasks.data(), n, clamps.data(), rec.data());
    if (rc) return rc;
    for (size_t k = 0; k < n; ++k) {                                                 // fill_resulBuf.v:377-378,422,429
        const bsw_aln_record& r = rec[k];
        uint32_t* o = rbb + 5 * k;
        o[0] = r.id;
        o[1] = ((uint32_t)r.qe << 16) | ((uint32_t)r.qb & 0xffffu);
        o[2] = ((uint32_t)r.re << 16) | ((uint32_t)r.rb & 0xffffu);
        o[3] = ((uint32_t)r.truesc << 16) | ((uint32_t)r.score & 0xffffu);
        o[4] = (uint32_t)r.w;
    }
    if (n_results) *n_results = (int)n;
    return BSW_OK;
}

}  // extern "C"
