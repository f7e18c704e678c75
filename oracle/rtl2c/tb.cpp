// Test bench around the cycle model that oracle/rtl2c/v2c.py generates from the mounted RTL (test infrastructure).
// Built into oracle/_ref/librtlsim.so by oracle/Makefile; loaded only by tests/ and tools/make_rtl_golden.py.
//
// Two harnesses, both driving the generated structs pin by pin:
//   rtl_sw_extend  one call of sw_pe_array_sw_extend (ports sw_pe_array_sw_extend.v:96-123), its base store being the
//                  reference's own query_mem module wired as sw_pe_array_proc_element.v:347-359,1237-1269 does
//   rtl_pe_array   one whole batch through sw_pe_array (task_parse -> 20 proc_element -> receive_match -> fill_resulBuf)
//                  between a TBB image and an RBB image; the two block RAMs either side are modelled as the hand-written
//                  shells use them: tbb.v:163-194 (synchronous read, data one clock after the address) and
//                  rbb.v:117-167 (one 32-bit word per write strobe)
#include "rtl_model.hpp"

#include <cstdint>
#include <cstring>
#include <memory>

using namespace rtl;

namespace {
using QueryMem = M_sw_pe_array_proc_element_query_mem_V__0;
using SwExtend = M_sw_pe_array_sw_extend;
using PeArray = M_sw_pe_array;

int16_t s16(uint64_t v) { return (int16_t)(uint16_t)v; }
}  // namespace

extern "C" {

// bases[0..nq) = query, bases[nq..nq+nt) = target (codes 0..4 in the low 3 bits of a nibble, as proc_element stores them).
// scalars: o_ins e_ins o_del e_del w h0 regScore max_ins max_del.  ret_in: the five *_ret_read carry-in ports
// (qle tle gtle gscore maxoff) -- proc_element passes its running state there (sw_pe_array_proc_element.v:361-396).
// out[7] = ap_return_0..6 = score, aw, qle, tle, gtle, gscore, max_off (sign-extended from 16 bit); returns the number of
// clocks from ap_start to ap_done, or -1 if the core never finished.
long rtl_sw_extend(const uint8_t *bases, int nq, int nt, const int *scalars, const int *ret_in, int *out,
                   uint64_t scramble_seed) {
    if (nq < 0 || nt < 0 || nq + nt > 2048 || nq > 255 || nt > 2047) return -2;
    auto mem = std::make_unique<QueryMem>();
    auto ext = std::make_unique<SwExtend>();
    if (scramble_seed) {
        uint64_t r = scramble_seed;
        mem->scramble(r);
        ext->scramble(r);
    }
    auto step = [&](bool wr, uint64_t waddr, uint64_t wdata) {
        settle(*mem);                                 // q0 is a register: it does not depend on this cycle's inputs
        ext->s_qs_V_q0 = mem->s_q0;
        settle(*ext);
        mem->s_we0 = wr;
        mem->s_ce0 = wr ? 1 : ext->s_qs_V_ce0;
        mem->s_address0 = wr ? waddr : ext->s_qs_V_address0;
        mem->s_d0 = wdata;
        settle(*mem);
    };
    auto tick = [&]() {
        mem->seq();
        ext->seq();
        mem->commit();
        ext->commit();
    };
    ext->s_ap_rst = 1;
    mem->s_reset = 1;
    for (int i = 0; i < 4; ++i) { step(false, 0, 0); tick(); }
    ext->s_ap_rst = 0;
    mem->s_reset = 0;
    for (int i = 0; i < nq + nt; ++i) { step(true, (uint64_t)i, bases[i] & 0xF); tick(); }

    ext->s_qs_baddr_V = 0;
    ext->s_ts_baddr_V = (uint64_t)nq;
    ext->s_qlen = (uint64_t)nq & 0xFF;
    ext->s_tlen_V = (uint64_t)nt & 0x7FF;
    ext->s_o_ins = scalars[0] & 0xFF;
    ext->s_e_ins = scalars[1] & 0xFF;
    ext->s_o_del = scalars[2] & 0xFF;
    ext->s_e_del = scalars[3] & 0xFF;
    ext->s_w_in = scalars[4] & 0xFF;
    ext->s_h0 = scalars[5] & 0xFF;
    ext->s_regScore_read = scalars[6] & 0xFFFF;
    ext->s_max_ins = scalars[7] & 0xFFFF;
    ext->s_max_del = scalars[8] & 0xFFFF;
    ext->s_qle_ret_read = ret_in[0] & 0xFFFF;
    ext->s_tle_ret_read = ret_in[1] & 0xFFFF;
    ext->s_gtle_ret_read = ret_in[2] & 0xFFFF;
    ext->s_gscore_ret_read = ret_in[3] & 0xFFFF;
    ext->s_maxoff_ret_read = ret_in[4] & 0xFFFF;
    ext->s_ap_start = 1;
    long clocks = 0;
    const long limit = 64L * 1024 * 1024;
    for (; clocks < limit; ++clocks) {
        step(false, 0, 0);
        if (ext->s_ap_done) {
            out[0] = s16(ext->s_ap_return_0);
            out[1] = s16(ext->s_ap_return_1);
            out[2] = s16(ext->s_ap_return_2);
            out[3] = s16(ext->s_ap_return_3);
            out[4] = s16(ext->s_ap_return_4);
            out[5] = s16(ext->s_ap_return_5);
            out[6] = s16(ext->s_ap_return_6);
            return clocks;
        }
        tick();
    }
    return -1;
}

// tbb: 65536 words (TBB image); rbb: 4096 words, pre-filled by the caller (the unused tail must come
// back untouched).  Returns clocks from the ap_start pulse to ap_done, -1 on timeout, -3 if the result writes were not the
// dense ascending sequence rbb.v's 16-word gluing assumes.  *n_words = number of result words written.
long rtl_pe_array(const uint32_t *tbb, uint32_t *rbb, int *n_words, long max_clocks, uint64_t scramble_seed) {
    auto top = std::make_unique<PeArray>();
    if (scramble_seed) {
        uint64_t r = scramble_seed;
        top->scramble(r);
    }
    uint32_t in_q = 0;                                   // tbb.v: RdDout one clock after RdAddr
    int written = 0;
    bool dense = true;
    auto cycle = [&]() {
        top->s_InData_q0 = in_q;
        settle(*top);
        uint32_t next_q = tbb[top->s_InData_address0 & 0xFFFF];
        if (top->s_ResData_we0) {
            uint64_t a = top->s_ResData_address0 & 0xFFF;
            if (a != (uint64_t)written) dense = false;
            rbb[a] = (uint32_t)top->s_ResData_d0;
            ++written;
        }
        bool done = top->s_ap_done != 0;
        top->seq();
        top->commit();
        in_q = next_q;
        return done;
    };
    top->s_ap_rst = 1;
    for (int i = 0; i < 8; ++i) cycle();
    top->s_ap_rst = 0;
    for (int i = 0; i < 4; ++i) cycle();
    written = 0;
    dense = true;
    top->s_ap_start = 1;                                 // tbb.v:123 task_start is a one-clock pulse
    bool done = cycle();
    top->s_ap_start = 0;
    long clocks = 1;
    while (!done && clocks < max_clocks) {
        done = cycle();
        ++clocks;
    }
    *n_words = written;
    if (!done) return -1;
    return dense ? clocks : -3;
}

}  // extern "C"
