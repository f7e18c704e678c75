// bsw_k3_core.cuh -- K3: one fused seed task per lane = what one FPGA processing element does
// (sw_pe_array_proc_element.v:1593-1685): left extension, right extension seeded with the left score, the 2-try band
// doubling that the RTL folds into sw_extend (sw_pe_array_sw_extend.v:1963,1824-1825,1969-1970), the local / to-end clip
// decision (pe:1672-1675) and the 5-word result record (pe:1187-1205,1662-1665).  The extensions themselves are
// k1_task (bsw_k1_core.cuh), so every score is bit-identical to the level-1 path.  __host__ __device__: csrc/emu.cpp runs
// the same code on the CPU for tests.
#pragma once
#include "bsw_k1_core.cuh"

namespace bsw {

struct SeedRecord { uint32_t id; int32_t qb, qe, rb, re, score, truesc, w; };     // == bsw_aln_record

// ksw_extend2's band clamp, evaluated exactly as BWA does (double division, truncation), or the wire's max_ins/max_del.
BSW_HD int k3_clamp(const DevParams& P, int qlen, int aw, int end_bonus, int max_ins_w, int max_del_w)
{
    int max_ins = max_ins_w, max_del = max_del_w;
    if (max_ins_w < 0) {
        max_ins = (int)((double)(qlen * P.max_mat + end_bonus - P.o_ins) / P.e_ins + 1.);
        max_ins = max_ins > 1 ? max_ins : 1;
        max_del = (int)((double)(qlen * P.max_mat + end_bonus - P.o_del) / P.e_del + 1.);
        max_del = max_del > 1 ? max_del : 1;
    }
    int w = aw < max_ins ? aw : max_ins;
    w = w < max_del ? w : max_del;
    return w;
}

constexpr int K3_MAX_BAND_TRY = 2;

// The RTL initialises max / max_i / max_j / max_ie / gscore once per sw_extend invocation, BEFORE its band-try loop
// (sw_pe_array_sw_extend.v:885-890,913-930,957-959), where ksw_extend2 starts every call afresh: the FPGA's second try
// continues from the first try's maxima (seen on ~2 % of the second tries of the translated
// RTL, tests/golden/rtl_*.npz).  Both update rules are monotone -- a row replaces (max, max_i, max_j) only when it beats
// the running max (sx:1959), and (gscore, max_ie) when it is not below the running gscore (sx:1941) -- so the carried
// second try equals this merge of the two stand-alone results.  Applied to wire tasks only: everything else follows BWA.
BSW_HD void k3_rtl_carry(const SlotResult& first, SlotResult& second)
{
    if (!(second.score > first.score)) { second.score = first.score; second.qle = first.qle; second.tle = first.tle; }
    if (second.gscore < first.gscore) { second.gscore = first.gscore; second.gtle = first.gtle; }
}

// spL/spR: the left / right flank (qlen == 0: no extension on that side, pe:1670).  The left flanks arrive reversed.
template <int VARIANT, int GENERIC, int SYM>
BSW_HD void k3_seed(const DevParams& P, int w, int pen_clip5, int pen_clip3, const SlotParam& spL, const SlotParam& spR,
                    const SeedParam& sd, int nqwL, int nqwR, uint32_t* eh, uint32_t* qsL, uint32_t* qsR,
                    const uint32_t* tgL, const uint32_t* tgR, SeedRecord& rec, uint32_t& cells)
{
    // initial state: pe:471-475,581-583,605-607,649-651,673-675,707-709,717-719,757-759,783-797
    int qb = 0, rb = 0, qe = spR.qlen, re = 0, score = 0;
    int sc0 = sd.init_score, truesc = sd.init_score;
    int aw0 = w, aw1 = w;
    cells = 0;
    for (int side = 0; side < 2; ++side) {                                       // pe:1597,1622
        const SlotParam& sp = side ? spR : spL;
        if (sp.qlen <= 0) continue;                                              // pe:1670,430,443-445
        const int h0 = side ? sc0 : sd.h0;                                       // pe:1671,1652
        const int pen_clip = side ? pen_clip3 : pen_clip5;
        int a_score = sc0, aw = w;
        SlotResult r;
        r.score = 0; r.qle = 0; r.tle = 0; r.gtle = 0; r.gscore = -1; r.max_off = 0; r.cells = 0; r.status = 0;
        for (int k = 0; k < K3_MAX_BAND_TRY; ++k) {                              // sx:1963,1878
            const int prev = a_score;                                            // sx:1822,1859
            aw = w << k;                                                         // sx:1765
            const int weff = k3_clamp(P, sp.qlen, aw, pen_clip, sd.max_ins[side], sd.max_del[side]);
            const SlotResult r0 = r;
            k1_task<VARIANT, GENERIC, SYM>(P, sp.qlen, sp.tlen, h0, weff, side ? nqwR : nqwL, eh, side ? qsR : qsL,
                                           side ? tgR : tgL, r, k == 0);
            cells += (uint32_t)r.cells;
            if (k > 0 && sd.max_ins[side] >= 0) k3_rtl_carry(r0, r);             // wire tasks only (bsw_fpga_batch)
            a_score = r.score;
            if (a_score == prev || r.max_off < (aw >> 1) + (aw >> 2)) break;     // sx:1824-1825,1969-1970,1837
        }
        if (side) aw1 = aw; else aw0 = aw;
        if (r.gscore <= 0 || r.gscore <= a_score - pen_clip) {                   // local: pe:1672,1674-1675,1667
            if (side == 0) { qb = sd.qbeg - r.qle; rb = -r.tle; truesc = a_score; }              // pe:591-599,659-667,767-777
            else           { qe = r.qle; re = r.tle; truesc += a_score - sc0; }                  // pe:615-623,683-691,1679-1680
        } else {                                                                 // to-end
            if (side == 0) { qb = 0; rb = -r.gtle; truesc = r.gscore; }
            else           { qe = spR.qlen; re = r.gtle; truesc += r.gscore - sc0; }
        }
        score = sc0 = a_score;                                                   // pe:697-700,727-728,1594,1685
    }
    rec.id = sd.id; rec.qb = qb; rec.qe = qe; rec.rb = rb; rec.re = re;          // pe:1187-1205,1662-1665
    rec.score = score; rec.truesc = truesc; rec.w = aw0 > aw1 ? aw0 : aw1;       // pe:1669,1684
}

}  // namespace bsw
