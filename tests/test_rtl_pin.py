"""The oracle -- and through it the CUDA path -- pinned to THE REFERENCE ITSELF.

tests/golden/rtl_*.npz hold outputs of the mounted Verilog (/root/reference/sw_pe_array*.v), produced by the cycle model
that oracle/rtl2c/v2c.py translates mechanically from those files (tools/make_rtl_golden.py; `make -C oracle`).  Nothing
in those files was computed by the C oracle or by the kernels.

  rtl_sw_extend.npz   6000 calls of sw_pe_array_sw_extend inside the RTL's numeric envelope (8-bit scores: h0 + qlen <= 127,
                      qlen <= 120, w <= 63, matrix +1/-4/-1, no z-drop): the 7 return values sw_pe_array_sw_extend.v:117-123
  rtl_pe_array.npz    8 whole batches through sw_pe_array: TBB image in, RBB image out
  rtl_sw_extend_wide.npz  3000 calls that leave the envelope (see test_rtl_width_model.py)

CPU suite: the C oracle (int32) must reproduce every in-envelope value; where the translated RTL is present
(oracle/_ref/librtlsim.so -- this container and the GPU box) a fresh random batch is simulated and compared as well, so
any disagreement between oracle/ksw_extend_ref.c and the RTL on an in-envelope task fails the suite.
GPU suite: the CUDA path through the C ABI must reproduce the same files.
"""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import assert_same, oracle_params, parse_tbb

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F6 = ("score", "qle", "tle", "gtle", "gscore", "max_off")


def load_l1(name="rtl_sw_extend.npz"):
    with np.load(os.path.join(GOLD, name)) as z:
        g = {k: z[k] for k in z.files}                                      # NpzFile re-reads an array on every access
    s = {k: g["scalars"][:, i] for i, k in enumerate(g["scalar_names"])}
    return g, s


def task_seqs(g, i):
    return g["qbuf"][g["qoff"][i]:g["qoff"][i + 1]], g["tbuf"][g["toff"][i]:g["toff"][i + 1]]


def test_oracle_reproduces_every_rtl_sw_extend_value(O):
    """Whole invocations (clamp from the host's max_ins/max_del, two band tries, maxima carried): all 7 returns."""
    g, s = load_l1()
    n = len(g["rtl"])
    assert n >= 6000 and (g["rtl"][:, 1] != s["w"]).sum() > 500          # the second try ran on many of them
    got = np.zeros_like(g["rtl"])
    cache = {}
    for i in range(n):
        q, t = task_seqs(g, i)
        assert O.rtl_envelope(len(q), len(t), int(s["h0"][i]), int(s["w"][i]), int(s["o_del"][i]), int(s["e_del"][i]), int(s["o_ins"][i]), int(s["e_ins"][i]))
        key = (int(s["o_del"][i]), int(s["e_del"][i]), int(s["o_ins"][i]), int(s["e_ins"][i]))
        if key not in cache:
            cache[key] = O.make_params(o_del=key[0], e_del=key[1], o_ins=key[2], e_ins=key[3], zdrop=0)
        p = cache[key]
        got[i], _ = O.sw_extend_rtl(p, q, t, int(s["h0"][i]), int(s["w"][i]), int(s["reg_score"][i]), int(s["max_ins"][i]), int(s["max_del"][i]))
    bad = np.nonzero((got != g["rtl"]).any(axis=1))[0]
    assert len(bad) == 0, f"{len(bad)} of {n} differ, first {bad[0]}: oracle {got[bad[0]]} rtl {g['rtl'][bad[0]]}"


def single_try_groups(g, s):
    """Tasks on which the RTL ran one band try, grouped by the per-batch scalars of the level-1 API."""
    one = np.nonzero(g["rtl"][:, 1] == s["w"])[0]
    keys = {}
    for i in one:
        keys.setdefault((int(s["o_del"][i]), int(s["e_del"][i]), int(s["o_ins"][i]), int(s["e_ins"][i]), int(s["end_bonus"][i])), []).append(i)
    return keys


def flat_subset(g, idx):
    qs, ts = zip(*(task_seqs(g, i) for i in idx))
    from helpers import flat_from_lists
    return flat_from_lists(qs, ts)


def test_oracle_ksw_extend2_entry_equals_rtl_on_single_try_tasks(O):
    """The entry point everything else is checked against (one ksw_extend2 call, band clamp from end_bonus)."""
    g, s = load_l1()
    total = 0
    for (o_del, e_del, o_ins, e_ins, eb), idx in single_try_groups(g, s).items():
        p = O.make_params(o_del=o_del, e_del=e_del, o_ins=o_ins, e_ins=e_ins, zdrop=0, end_bonus=eb)
        qbuf, qoff, tbuf, toff = flat_subset(g, idx)
        res, _ = O.extend_batch(p, qbuf, qoff, tbuf, toff, s["h0"][idx], s["w"][idx])
        for k, f in zip((0, 2, 3, 4, 5, 6), F6):
            assert_same(res[f], g["rtl"][idx, k], f"{f} (gaps {o_del},{e_del},{o_ins},{e_ins} end_bonus {eb})")
        total += len(idx)
    assert total > 4000


def golden_batches():
    g = np.load(os.path.join(GOLD, "rtl_pe_array.npz"))
    return [(g["tbb"][b], g["rbb"][b], int(g["n_words"][b])) for b in range(len(g["tbb"]))]


def seeds_of(tbb):
    hdr, tasks = parse_tbb(tbb)
    seeds, gaps = [], []
    for t in tasks:
        b = np.array(t["bases"], np.uint8)
        ql, qr, tl, tr = t["qlen"][0], t["qlen"][1], t["tlen"][0], t["tlen"][1]
        init = t["init_score"] - 65536 if t["init_score"] >= 32768 else t["init_score"]
        seeds.append(dict(q_left=b[:ql], q_right=b[ql:ql + qr], t_left=b[ql + qr:ql + qr + tl], t_right=b[ql + qr + tl:],
                          init_score=init, qbeg=t["qbeg"], h0=t["h0"], id=t["id"]))
        gaps.append([t["max_ins"][0], t["max_del"][0], t["max_ins"][1], t["max_del"][1]])
    return hdr, seeds, np.array(gaps, np.int32).reshape(len(seeds), 4)


def records_of(B, rbb, n):
    r = B.rbb_decode(rbb, n)
    return r[np.argsort(r["id"], kind="stable")]


def params2_of(mod, hdr):
    h = {k: int(v) for k, v in hdr.items()}
    return mod.make_params2(mod.make_params(o_del=h["o_del"], e_del=h["e_del"], o_ins=h["o_ins"], e_ins=h["e_ins"], zdrop=0),
                            w=h["w"], pen_clip5=h["pen_clip5"], pen_clip3=h["pen_clip3"])


def test_oracle_reproduces_rtl_batches(O, B):
    """task_parse -> 20 proc_element -> receive_match -> fill_resulBuf: the records, their number, the untouched tail."""
    second = carried = 0
    for tbb, rbb, nw in golden_batches():
        hdr, seeds, gaps = seeds_of(tbb)
        n = len(seeds)
        assert nw == 5 * n and (rbb[nw:] == 0xDEADBEEF).all()              # dense from word 0, nothing after (fill_resulBuf.v:377-429)
        if n == 0:
            continue
        tasks, keep = B.make_seed_tasks(seeds)
        ot = (O.SeedTask * n)()
        C.memmove(ot, tasks, C.sizeof(tasks))
        want, _ = O.chain2aln_rtl(oracle_params(O, params2_of(B, hdr)), ot, gaps)
        got = records_of(B, rbb, n)
        assert_same(got, want[np.argsort(want["id"], kind="stable")], "RBB records")
        second += int((got["w"] != hdr["w"]).sum())
        bwa, _ = O.chain2aln_batch(oracle_params(O, params2_of(B, hdr)), ot)
        carried += int((bwa[np.argsort(bwa["id"], kind="stable")] != got).sum())
    assert second > 400                                                    # band retries are covered ...
    assert carried > 20                                                    # ... including records on which ksw_extend2's fresh
    #                                                                        second try and the FPGA's carried one differ


def test_product_encoder_builds_the_image_the_rtl_consumed(B):
    """bsw_tbb_encode against the images tools/make_rtl_golden.py wrote from the RTL's field map and fed to the RTL."""
    for tbb, _, _ in golden_batches():
        hdr, seeds, _ = seeds_of(tbb)
        if not seeds:
            continue
        mine = B.tbb_encode(params2_of(B, hdr), seeds)
        n = len(seeds)
        a, b = tbb.copy(), mine.copy()
        a[10:8 + 8 * n:8] -= a[10]                                         # data offsets are relative to task 0's (task_parse.v:1928)
        b[10:8 + 8 * n:8] -= b[10]
        used = 8 + 8 * n + int(a[10 + 8 * (n - 1)]) + (sum(len(seeds[-1][k]) for k in ("q_left", "q_right", "t_left", "t_right")) + 7) // 8
        assert np.array_equal(a[:3], b[:3]) and np.array_equal(a[8:used], b[8:used])


def test_emulated_fused_seed_kernel_equals_rtl_batches(B):
    """The K3 lane function compiled for the host, in the wire mode bsw_fpga_batch uses (host-supplied clamp, carried
    second try), against the RTL's records."""
    for tbb, rbb, nw in golden_batches():
        hdr, seeds, gaps = seeds_of(tbb)
        if not seeds:
            continue
        got = B.emu_chain2aln(params2_of(B, hdr), seeds, wire_gaps=gaps)
        assert_same(got[np.argsort(got["id"], kind="stable")], records_of(B, rbb, len(seeds)), "emulated K3 vs RTL")


def test_live_translated_rtl_agrees_with_oracle(O):
    """A batch the golden files do not contain, simulated now (only where oracle/_ref/librtlsim.so exists)."""
    from oracle import rtlsim as R
    if not R.available():
        pytest.skip("oracle/_ref/librtlsim.so not built (needs /root/reference)")
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLD), "..", "tools"))
    import make_rtl_golden as G
    tasks = G.level1_tasks(np.random.default_rng(int.from_bytes(os.urandom(4), "little")), 300, wide=False)
    for i, t in enumerate(tasks):
        rtl, _ = R.sw_extend(t["q"], t["t"], t["h0"], t["w"], t["o_ins"], t["e_ins"], t["o_del"], t["e_del"], t["reg_score"],
                             t["max_ins"], t["max_del"], scramble_seed=1000 + i)
        p = O.make_params(o_del=t["o_del"], e_del=t["e_del"], o_ins=t["o_ins"], e_ins=t["e_ins"], zdrop=0)
        ora, _ = O.sw_extend_rtl(p, t["q"], t["t"], t["h0"], t["w"], t["reg_score"], t["max_ins"], t["max_del"])
        assert np.array_equal(rtl, ora), f"task {i}: rtl {rtl} oracle {ora} ({t})"


# ---------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_cuda_level1_equals_rtl(B, ctx):
    g, s = load_l1()
    total = 0
    for (o_del, e_del, o_ins, e_ins, eb), idx in single_try_groups(g, s).items():
        p = B.make_params(o_del=o_del, e_del=e_del, o_ins=o_ins, e_ins=e_ins, zdrop=0, end_bonus=eb)
        qbuf, qoff, tbuf, toff = flat_subset(g, idx)
        res, _ = ctx.sw_extend_batch(p, qbuf, qoff, tbuf, toff, s["h0"][idx], s["w"][idx])
        for k, f in zip((0, 2, 3, 4, 5, 6), F6):
            assert_same(res[f], g["rtl"][idx, k], f"CUDA {f} vs RTL")
        total += len(idx)
    assert total > 4000


@pytest.mark.gpu
def test_cuda_fpga_batch_equals_rtl(B, ctx):
    """bsw_fpga_batch: same TBB image in, same records out as the translated sw_pe_array (record order aside: the RTL
    emits in completion order, identified by word 0), nothing written past them."""
    for tbb, rbb, nw in golden_batches():
        n = nw // 5
        mine, nres = ctx.pe_array_batch(tbb)
        assert nres == n
        assert (mine[nw:] == 0).all()
        if n:
            assert_same(records_of(B, mine, n), records_of(B, rbb, n), "bsw_fpga_batch vs RTL")


@pytest.mark.gpu
def test_fpga_strict_refuses_what_the_fpga_cannot_compute(B, ctx):
    """Option fpga_strict: a TBB with a task outside the 8-bit envelope is refused (BSW_ERANGE) instead of answered with
    ksw_extend2's int32 result, which the FPGA would not return."""
    q = np.zeros(60, np.uint8)
    P2 = B.make_params2(B.make_params(zdrop=0), w=50, pen_clip5=5, pen_clip3=5)
    ok = dict(q_left=q[:20], q_right=q[:30], t_left=q[:25], t_right=q[:35], init_score=19, qbeg=20, h0=19, id=1)
    hot = dict(ok, h0=120, init_score=120, id=2)
    ctx.set_option("fpga_strict", 1)
    try:
        rbb, n = ctx.pe_array_batch(B.tbb_encode(P2, [ok, ok]))
        assert n == 2
        with pytest.raises(B.BswError) as e:
            ctx.pe_array_batch(B.tbb_encode(P2, [ok, hot]))
        assert e.value.code == B.BSW_ERANGE and "index 1" in str(e.value)
    finally:
        ctx.set_option("fpga_strict", 0)
    rbb, n = ctx.pe_array_batch(B.tbb_encode(P2, [ok, hot]))               # default: answered (ksw_extend2 semantics)
    assert n == 2
