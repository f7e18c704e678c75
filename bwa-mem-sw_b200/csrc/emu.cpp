// emu.cpp -- TEST INFRASTRUCTURE ONLY (libbsw_emu.so).  NOT linked into libbsw.so, never used by the product path.
//
// Runs the host half of the level-1 call exactly as libbsw.so does (validate -> plan -> pack into the tiled arena)
// and then executes the K1 lane function (bsw_k1_core.cuh, the same source the device kernel instantiates) on the
// CPU, lane by lane, over emulated shared-memory arrays.  This lets `pytest -m "not gpu"` check the scheduler, the
// packer and the K1 control flow (lazy narrowing, partial chunks, DPX identities) against the oracle without a GPU.
// The CUDA kernels themselves are only ever checked on a GPU (`pytest -m gpu`).
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bsw.h"
#include "bsw_device.cuh"
#include "bsw_k1_core.cuh"
#include "bsw_sched.h"

using namespace bsw;

namespace {

template <int VARIANT, int GENERIC, int SYM>
void run_tile(const DevParams& dp, const TileHdr& hd, const SlotParam* slots, const uint32_t* arena, SlotResult* out,
              int qmax, int nqw_max)
{
    const int nqw = (int)(hd.nqw_ntw & 0xffffu);
    std::vector<uint32_t> qs((size_t)(nqw_max + 1) * K1_S, 0xdeadbeefu);
    std::vector<uint32_t> eh((size_t)(qmax + 1 + K1_EH_SLACK) * K1_S, 0xdeadbeefu);
    memcpy(qs.data(), arena + (size_t)hd.qoff16 * 4, (size_t)nqw * K1_S * 4);       // the TMA bulk copy
    for (int lane = 0; lane < K1_S; ++lane) qs[(size_t)nqw * K1_S + lane] = 0;
    for (int lane = 0; lane < K1_S; ++lane) {
        const SlotParam& sp = slots[hd.slot0 + lane];
        if (sp.qlen <= 0) continue;
        const uint32_t* tg = arena + (size_t)hd.toff16 * 4 + lane;
        k1_task<VARIANT, GENERIC, SYM>(dp, sp.qlen, sp.tlen, sp.h0, sp.w, eh.data() + lane, qs.data() + lane, tg,
                                       out[hd.slot0 + lane]);
    }
}

}  // namespace

extern "C" {

// info[0] = launches, info[1] = tiles, info[2] = arena words, info[3] = padded lanes
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
    opt.variant = variant; opt.force_kernel = 1; opt.fast_matrix = fast; opt.host_threads = 4;
    std::vector<ExtTask> v(n);
    for (size_t i = 0; i < n; ++i) {
        ExtTask& x = v[i];
        x.q = qbuf + qoff[i]; x.t = tbuf + toff[i];
        x.qlen = (int32_t)(qoff[i + 1] - qoff[i]); x.tlen = (int32_t)(toff[i + 1] - toff[i]); x.h0 = h0[i];
        x.w = (x.qlen >= 1 && w[i] >= 0) ? clamp_band(params->mat, x.qlen, w[i], params->end_bonus, params->o_ins,
                                                      params->e_ins, params->o_del, params->e_del) : -1;
    }
    std::vector<uint8_t> cls(n);
    size_t bad = 0; std::string msg;
    int rc = validate_tasks(v.data(), n, mx, opt, cls.data(), &bad, &msg);
    if (rc) return rc;
    Plan P;
    build_plan(v.data(), cls.data(), n, opt, &P);
    std::vector<uint32_t> arena(P.arena_words + 64, 0xdeadbeefu);
    pack_arena(v.data(), P, opt, arena.data());
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
    if (info) { info[0] = (int64_t)P.launches.size(); info[1] = (int64_t)P.tiles.size(); info[2] = (int64_t)P.arena_words; info[3] = (int64_t)pad; }
    return BSW_OK;
}

}  // extern "C"
