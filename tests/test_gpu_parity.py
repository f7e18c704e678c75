"""Parity tests proper: the CUDA path, called through the C ABI (libbsw.so), against the CPU oracle on the same
seeded inputs and against the frozen golden vectors.  All outputs are integers: the bar is bit-exact."""
import glob
import os

import numpy as np
import pytest

from helpers import assert_same, flat_from_lists, oracle_chain2aln, seeds_from_flat

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
L1 = sorted(glob.glob(os.path.join(GOLD, "l1_*.npz")))


def both(B, O, ctx, t, variant=1, opts=None, **pk):
    p, po = B.make_params(**pk), O.make_params(**pk)
    ctx.set_option("variant", variant)
    for k, v in (opts or {}).items():
        ctx.set_option(k, v)
    try:
        ro, co = O.extend_batch(po, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"], variant=variant)
        rg, cg = ctx.sw_extend_batch(p, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    finally:
        ctx.set_option("variant", 1); ctx.set_option("force_kernel", 0); ctx.set_option("k2_min_qlen", 384); ctx.set_option("wide", 1)
    assert_same(ro, rg, "results")
    assert_same(co.astype(np.int64), cg.astype(np.int64), "cells")
    return ro, co


def test_native_library_is_loaded(B, ctx):
    maps = open("/proc/self/maps").read()
    assert "libbsw.so" in maps
    assert ctx.num_devices >= 1


@pytest.mark.parametrize("name,n", [("cfg1_101bp", 100_000), ("cfg2_150bp", 60_000), ("cfg3_mixed", 60_000)])
def test_k1_workloads(B, O, ctx, name, n):
    both(B, O, ctx, B.synth_tasks(name, n, seed=2))


@pytest.mark.parametrize("variant", [1, 2])
def test_k1_variants_with_ambiguous_bases(B, O, ctx, variant):
    both(B, O, ctx, B.synth_tasks("cfg3_mixed", 20_000, seed=3, n_frac=0.01), variant=variant)


@pytest.mark.parametrize("pk", [dict(o_del=4, e_del=2, o_ins=7, e_ins=1), dict(a=2, b=3, zdrop=20),
                                dict(zdrop=0, o_del=0, o_ins=0), dict(o_del=0, e_del=3, o_ins=9, e_ins=2, end_bonus=0)])
def test_k1_scoring_variations(B, O, ctx, pk):
    both(B, O, ctx, B.synth_tasks("cfg3_mixed", 10_000, seed=4), **pk)
    both(B, O, ctx, B.synth_tasks("cfg3_mixed", 10_000, seed=4), variant=2, **pk)


def test_custom_matrix(B, O, ctx):
    m = B.bwa_fill_scmat(1, 4).copy()
    m[1], m[7], m[24] = 2, -3, 0
    both(B, O, ctx, B.synth_tasks("cfg3_mixed", 10_000, seed=5), mat=m)
    both(B, O, ctx, B.synth_tasks("cfg3_mixed", 3_000, seed=5), mat=m, opts={"force_kernel": 2})


def test_k2_on_short_tasks(B, O, ctx):
    """The intra-task kernel forced onto short/mixed tasks: same answers as K1 and the oracle."""
    both(B, O, ctx, B.synth_tasks("cfg3_mixed", 10_000, seed=6), opts={"force_kernel": 2})
    both(B, O, ctx, B.synth_tasks("cfg3_mixed", 5_000, seed=6, n_frac=0.02), opts={"force_kernel": 2})
    both(B, O, ctx, B.synth_tasks("cfg3_mixed", 5_000, seed=7), opts={"force_kernel": 2}, o_del=4, e_del=2, o_ins=7, e_ins=1, zdrop=30)


def test_k2_long_reads(B, O, ctx):
    """BASELINE config 4: 1-10 kb, w=500, PacBio-like errors (zdrop 100 and 400)."""
    both(B, O, ctx, B.synth_tasks("cfg4_long", 96, seed=8))
    both(B, O, ctx, B.synth_tasks("cfg4_long", 48, seed=9), zdrop=400)


def test_long_batch_spreads_over_chunks(B, O, ctx):
    """Adaptive chunking: a batch of long tasks is cut by bases, not by task count, so several workers take part."""
    t = B.synth_tasks("cfg4_long", 1500, seed=90)
    ctx.reset_stats()
    both(B, O, ctx, t)
    assert ctx.stats()["kernel_launches"] >= 2          # more than one chunk


def test_mixed_short_and_long_in_one_batch(B, O, ctx):
    a, b = B.synth_tasks("cfg3_mixed", 3000, seed=10), B.synth_tasks("cfg4_long", 8, seed=10)
    t = dict(qbuf=np.concatenate([a["qbuf"][:a["qoff"][-1]], b["qbuf"]]), tbuf=np.concatenate([a["tbuf"][:a["toff"][-1]], b["tbuf"]]),
             qoff=np.concatenate([a["qoff"], b["qoff"][1:] + a["qoff"][-1]]), toff=np.concatenate([a["toff"], b["toff"][1:] + a["toff"][-1]]),
             h0=np.concatenate([a["h0"], b["h0"]]), w=np.concatenate([a["w"], b["w"]]))
    both(B, O, ctx, t)
    both(B, O, ctx, t, opts={"k2_min_qlen": 100})


def test_edge_shapes(B, O, ctx):
    rng = np.random.default_rng(9)
    qs, ts, h0, w = [], [], [], []
    for qlen, tlen, h, ww in [(1, 1, 1, 100), (1, 50, 30, 100), (50, 1, 30, 100), (8, 8, 5, 0), (9, 200, 100, 3),
                              (16, 16, 1, 100), (17, 40, 300, 1), (255, 600, 19, 100), (300, 310, 1000, 50),
                              (256, 256, 20, 100), (257, 513, 20, 100), (1536, 1600, 50, 100), (1537, 1600, 50, 100)]:
        q = rng.integers(0, 4, qlen).astype(np.uint8)
        t = np.resize(q, tlen).astype(np.uint8)
        t[rng.random(tlen) < 0.1] = 3
        qs.append(q); ts.append(t); h0.append(h); w.append(ww)
    qbuf, qoff, tbuf, toff = flat_from_lists(qs, ts)
    t = dict(qbuf=qbuf, qoff=qoff, tbuf=tbuf, toff=toff, h0=np.array(h0, np.int32), w=np.array(w, np.int32))
    both(B, O, ctx, t)
    both(B, O, ctx, t, opts={"force_kernel": 2})
    both(B, O, ctx, t, variant=2)                                    # 1536 / 1537: the K1 / K2 border under V2 as well
    both(B, O, ctx, t, variant=2, opts={"force_kernel": 2})


def test_v2_on_the_intra_task_kernel(B, O, ctx):
    """Upstream-BWA semantics (V2) on K2: the zero guard, gap opens from M, the conditional first column and the
    zero-scan narrowing with its stale row-buffer slots -- short tasks forced onto K2, then 1-10 kb reads (ring row
    buffer), then other scorings and N bases."""
    both(B, O, ctx, B.synth_tasks("cfg3_mixed", 20_000, seed=6), variant=2, opts={"force_kernel": 2})
    both(B, O, ctx, B.synth_tasks("cfg3_mixed", 5_000, seed=7, n_frac=0.01), variant=2, opts={"force_kernel": 2})
    for zd in (100, 0):
        both(B, O, ctx, B.synth_tasks("cfg4_long", 1500, seed=8), variant=2, zdrop=zd)
    both(B, O, ctx, B.synth_tasks("cfg4_long", 300, seed=9), variant=2, o_del=4, e_del=2, o_ins=7, e_ins=1)
    both(B, O, ctx, B.synth_tasks("cfg3_mixed", 5_000, seed=10), variant=2, opts={"force_kernel": 2}, a=2, b=3, zdrop=20)
    from helpers import random_small_tasks
    both(B, O, ctx, random_small_tasks(np.random.default_rng(11), 6000), variant=2, opts={"force_kernel": 2})


def test_empty_batch_and_errors(B, ctx):
    p = B.make_params()
    out, _ = ctx.sw_extend_batch(p, np.zeros(8, np.uint8), np.zeros(1, np.int64), np.zeros(8, np.uint8), np.zeros(1, np.int64),
                                 np.zeros(0, np.int32), np.zeros(0, np.int32))
    assert len(out) == 0
    q = np.array([0, 1, 9, 2], np.uint8)
    qbuf, qoff, tbuf, toff = flat_from_lists([q], [q])
    with pytest.raises(B.BswError) as e:
        ctx.sw_extend_batch(p, qbuf, qoff, tbuf, toff, [10], [100])
    assert e.value.code == B.BSW_EINVAL and "task 0" in str(e.value)
    good = np.array([0, 1, 2, 3], np.uint8)
    qbuf, qoff, tbuf, toff = flat_from_lists([good], [good])
    with pytest.raises(B.BswError) as e:                     # h0 must be > 0 (sw_extend precondition)
        ctx.sw_extend_batch(p, qbuf, qoff, tbuf, toff, [0], [100])
    assert e.value.code == B.BSW_EINVAL
    ctx.set_option("wide", 0)
    try:
        with pytest.raises(B.BswError) as e:                 # 16-bit row-state envelope, with the 32-bit kernel switched off
            ctx.sw_extend_batch(p, qbuf, qoff, tbuf, toff, [40000], [100])
        assert e.value.code == B.BSW_ERANGE
    finally:
        ctx.set_option("wide", 1)
    with pytest.raises(B.BswError) as e:                     # beyond the 32-bit kernel's envelope as well
        ctx.sw_extend_batch(p, qbuf, qoff, tbuf, toff, [0x7fffff00], [100])
    assert e.value.code == B.BSW_ERANGE
    bad = B.make_params(e_del=0)
    with pytest.raises(B.BswError):
        ctx.sw_extend_batch(bad, qbuf, qoff, tbuf, toff, [10], [100])


def test_task_record_layout_and_async(B, O, ctx):
    t = B.synth_tasks("cfg3_mixed", 500, seed=12)
    qs = [t["qbuf"][t["qoff"][i]:t["qoff"][i + 1]] for i in range(500)]
    ts = [t["tbuf"][t["toff"][i]:t["toff"][i + 1]] for i in range(500)]
    p, po = B.make_params(), O.make_params()
    ro, _ = O.extend_batch(po, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    assert_same(ro, ctx.sw_extend_tasks(p, qs, ts, t["h0"], t["w"]), "bsw_extend_batch")
    # async pair: submit / poll / wait (the FPGA's "write REQ_PEARRAY, poll the DSM busy bit")
    import ctypes as C
    qa = [np.ascontiguousarray(q) for q in qs]; ta = [np.ascontiguousarray(x) for x in ts]
    tasks = (B.Task * 500)()
    for i in range(500):
        tasks[i].query, tasks[i].target = qa[i].ctypes.data, ta[i].ctypes.data
        tasks[i].qlen, tasks[i].tlen, tasks[i].h0, tasks[i].w = len(qa[i]), len(ta[i]), int(t["h0"][i]), int(t["w"][i])
    out = np.zeros(500, dtype=B.RESULT_DTYPE)
    tk = ctx.submit(p, tasks, 500, out)
    ctx.wait(tk)
    assert_same(ro, out, "async")
    # ticket life: submit -> poll until 1 -> wait (poll never releases the ticket; a second wait is refused)
    out[:] = 0
    tk = ctx.submit(p, tasks, 500, out)
    import time
    t0 = time.time()
    while ctx.poll(tk) == 0 and time.time() - t0 < 30:
        time.sleep(0.0005)
    assert ctx.poll(tk) == 1 and ctx.poll(tk) == 1
    ctx.wait(tk)
    assert_same(ro, out, "async after polling")
    with pytest.raises(B.BswError):
        ctx.wait(tk)
    # four tickets open at once (the FPGA's four PE arrays), each on its own slice
    outs = [np.zeros(125, dtype=B.RESULT_DTYPE) for _ in range(4)]
    tks = [ctx.submit(p, C.cast(C.byref(tasks, 125 * k * C.sizeof(B.Task)), C.POINTER(B.Task)), 125, outs[k]) for k in range(4)]
    for tk in tks:
        ctx.wait(tk)
    assert_same(ro, np.concatenate(outs), "four async batches in flight")


def test_four_threads_share_one_context(B, O, ctx):
    """Concurrent blocking calls on ONE context overlap (no context-wide lock: every call takes its own set of worker
    pipelines) and stay bit-exact; level 1, level 2 and level 3 mixed.  Each thread sees its own error text."""
    import threading
    t = B.synth_tasks("cfg3_mixed", 60_000, seed=90)
    p, po = B.make_params(), O.make_params()
    flat = (t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    ro, _ = O.extend_batch(po, *flat)
    seeds = seeds_from_flat(B.synth_tasks("cfg1_101bp", 1600, seed=91), 800, unset_score_every=4)
    P2 = B.make_params2(B.make_params(zdrop=0), w=100, pen_clip5=5, pen_clip3=5)
    want2, _ = oracle_chain2aln(O, B, P2, seeds)
    tbb = B.tbb_encode(P2, seeds)
    errs = []

    def level1(k):
        for _ in range(4):
            r, _ = ctx.sw_extend_batch(p, *flat)
            if not np.array_equal(r, ro):
                errs.append(f"thread {k}: level 1 differs")

    def level2(k):
        for _ in range(6):
            if not np.array_equal(ctx.proc_element_batch(P2, seeds), want2):
                errs.append(f"thread {k}: level 2 differs")

    def level3(k):
        for _ in range(10):
            rbb, n = ctx.pe_array_batch(tbb)
            if n != 800 or not np.array_equal(B.rbb_decode(rbb, n), want2):
                errs.append(f"thread {k}: level 3 differs")

    def failing(k):
        bad = t["h0"].copy(); bad[17] = 0
        for _ in range(4):
            try:
                ctx.sw_extend_batch(p, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], bad, t["w"])
                errs.append("an invalid batch was accepted")
            except B.BswError as e:
                if "task 17" not in str(e):
                    errs.append(f"wrong error text in the failing thread: {e}")

    th = [threading.Thread(target=f, args=(k,)) for k, f in enumerate((level1, level2, level3, failing, level1))]
    for x in th:
        x.start()
    for x in th:
        x.join()
    assert not errs, errs


@pytest.mark.parametrize("path", L1, ids=[os.path.basename(p) for p in L1])
def test_golden_level1(B, ctx, path):
    g = np.load(path)
    s = g["scal"]
    p = B.make_params(mat=g["mat"], o_del=int(s[0]), e_del=int(s[1]), o_ins=int(s[2]), e_ins=int(s[3]), zdrop=int(s[4]), end_bonus=int(s[5]))
    ctx.set_option("variant", int(s[6]))
    try:
        res, cells = ctx.sw_extend_batch(p, g["qbuf"], g["qoff"], g["tbuf"], g["toff"], g["h0"], g["w"])
    finally:
        ctx.set_option("variant", 1)
    assert_same(res, g["res"], "CUDA vs golden")
    assert_same(cells.astype(np.int64), g["cells"], "cells")


def test_golden_level2(B, ctx):
    g = np.load(os.path.join(GOLD, "l2_cfg3.npz"))
    seeds = seeds_from_flat(g, int(g["nreads"]), unset_score_every=3)
    got = ctx.proc_element_batch(B.make_params2(w=100, pen_clip5=5, pen_clip3=5), seeds)
    assert_same(got, g["rec"], "level 2 vs golden")


@pytest.mark.parametrize("w,zdrop", [(100, 100), (10, 100), (5, 0)])
def test_level2_proc_element(B, O, ctx, w, zdrop):
    """left + right extension, band-doubling retry (small w forces the second try), clip decision."""
    t = B.synth_tasks("cfg3_mixed", 4000, seed=20 + w)
    seeds = seeds_from_flat(t, 2000, unset_score_every=3)
    P2 = B.make_params2(B.make_params(zdrop=zdrop), w=w, pen_clip5=5, pen_clip3=7)
    want, _ = oracle_chain2aln(O, B, P2, seeds)
    got = ctx.proc_element_batch(P2, seeds)
    assert_same(want, got, "proc_element")
    if w < 100:
        assert (want["w"] == 2 * w).any()          # the retry path ran


def test_level2_many_chunks_and_errors_found_by_the_workers(B, O, ctx):
    """30 k seeds = four pipeline chunks; validation runs inside the chunk workers, so a bad seed deep in the batch must
    still fail the call with its own index, and a bad base code with the flank it sits in."""
    t = B.synth_tasks("cfg2_150bp", 60000, seed=31)
    seeds = seeds_from_flat(t, 30000, unset_score_every=3)
    P2 = B.make_params2(B.make_params(), w=100, pen_clip5=5, pen_clip3=5)
    want, _ = oracle_chain2aln(O, B, P2, seeds)
    got = ctx.proc_element_batch(P2, seeds)
    assert_same(want, got, "30 k seeds")
    bad = dict(seeds[20000]); bad["h0"] = 0                          # a left flank with h0 < 1
    if len(bad["q_left"]) == 0:
        bad = dict(seeds[20001]); bad["h0"] = 0; k = 20001
    else:
        k = 20000
    broken = list(seeds); broken[k] = bad
    with pytest.raises(B.BswError) as e:
        ctx.proc_element_batch(P2, broken)
    assert e.value.code == B.BSW_EINVAL and f"seed task {k}" in str(e.value)
    k = next(i for i in range(25000, 30000) if len(seeds[i]["q_right"]) > 3)
    bad = dict(seeds[k]); qr = np.array(bad["q_right"], dtype=np.uint8).copy(); qr[2] = 7; bad["q_right"] = qr
    broken = list(seeds); broken[k] = bad
    with pytest.raises(B.BswError) as e:
        ctx.proc_element_batch(P2, broken)
    assert e.value.code == B.BSW_EINVAL and f"seed task {k}" in str(e.value) and "right flank" in str(e.value)
    assert_same(want, ctx.proc_element_batch(P2, seeds), "the context still works after the failed calls")


def test_registered_buffers_skip_the_staging_pass(B, O, ctx):
    """Raw mode: bases in registered host memory are copied in place and packed on the device.  Same results, per-task
    cells and error codes as the staged path; tasks with N are rerun with matrix lookup; long tasks fall back."""
    t = B.synth_tasks("cfg3_mixed", 40000, seed=41, n_frac=0.002)
    p, po = B.make_params(), O.make_params()
    flat = (t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    ro, co = O.extend_batch(po, *flat)
    ctx.register_host(t["qbuf"]); ctx.register_host(t["tbuf"])
    ctx.set_option("raw_inputs", 1)                  # default is "auto": raw only when host threads are scarce
    try:
        ctx.reset_stats()
        r, c = ctx.sw_extend_batch(p, *flat)
        st = ctx.stats()
        assert_same(ro, r, "raw mode results"); assert_same(co.astype(np.int64), c.astype(np.int64), "raw mode cells")
        assert st["h2d_bytes"] > int(t["qoff"][-1] + t["toff"][-1])            # the bases crossed PCIe one byte each
        for variant in (2,):
            ctx.set_option("variant", variant)
            try:
                r2, c2 = ctx.sw_extend_batch(p, *flat)
            finally:
                ctx.set_option("variant", 1)
            ro2, co2 = O.extend_batch(po, *flat, variant=variant)
            assert_same(ro2, r2, "raw mode V2"); assert_same(co2.astype(np.int64), c2.astype(np.int64), "raw mode V2 cells")
        pk = dict(o_del=4, e_del=2, o_ins=7, e_ins=1, a=2, b=3)
        r3, c3 = ctx.sw_extend_batch(B.make_params(**pk), *flat)
        ro3, co3 = O.extend_batch(O.make_params(**pk), *flat)
        assert_same(ro3, r3, "raw mode, other scoring")
        custom = B.bwa_fill_scmat(1, 4).copy(); custom[1] = -2; custom[5] = -2                     # not +a/-b: matrix-lookup kernel, no rerun
        r4, _ = ctx.sw_extend_batch(B.make_params(mat=custom), *flat)
        ro4, _ = O.extend_batch(O.make_params(mat=custom), *flat)
        assert_same(ro4, r4, "raw mode, custom matrix")
        bad = t["tbuf"][int(t["toff"][12345]) + 3]
        t["tbuf"][int(t["toff"][12345]) + 3] = 9
        with pytest.raises(B.BswError) as e:
            ctx.sw_extend_batch(p, *flat)
        t["tbuf"][int(t["toff"][12345]) + 3] = bad
        assert e.value.code == B.BSW_EINVAL and "task 12345" in str(e.value)
        ctx.set_option("raw_inputs", 0)
        ctx.reset_stats()
        r5, _ = ctx.sw_extend_batch(p, *flat)
        assert ctx.stats()["h2d_bytes"] < st["h2d_bytes"]                     # staged again: 4 bit per base
        ctx.set_option("raw_inputs", 1)
        assert_same(ro, r5, "staged path on registered buffers")
        # a mixed batch with long tasks: those chunks are staged, the rest stays raw
        lt = B.synth_tasks("cfg4_long", 6, seed=42)
        qbuf = np.concatenate([t["qbuf"][:t["qoff"][2000]], lt["qbuf"]]); tbuf = np.concatenate([t["tbuf"][:t["toff"][2000]], lt["tbuf"]])
        qoff = np.concatenate([t["qoff"][:2001], lt["qoff"][1:] + t["qoff"][2000]]); toff = np.concatenate([t["toff"][:2001], lt["toff"][1:] + t["toff"][2000]])
        h0 = np.concatenate([t["h0"][:2000], lt["h0"]]); w = np.concatenate([t["w"][:2000], lt["w"]])
        ctx.register_host(qbuf); ctx.register_host(tbuf)
        try:
            r6, c6 = ctx.sw_extend_batch(p, qbuf, qoff, tbuf, toff, h0, w)
        finally:
            ctx.unregister_host(qbuf); ctx.unregister_host(tbuf)
        ro6, co6 = O.extend_batch(po, qbuf, qoff, tbuf, toff, h0, w)
        assert_same(ro6, r6, "raw + long tasks")
        # level 2 right after raw-mode level-1 chunks reuses the same stream slots (regression: stale raw flag)
        seeds = seeds_from_flat(t, 3000, unset_score_every=3)
        P2 = B.make_params2(p, w=100, pen_clip5=5, pen_clip3=5)
        want2, _ = oracle_chain2aln(O, B, P2, seeds)
        assert_same(want2, ctx.proc_element_batch(P2, seeds), "level 2 after raw mode")
        # auto: with the device planner (default) every chunk goes raw or 2-bit packed depending on whether the copy engine
        # keeps up (timing dependent: results only); with the host planner raw only when host threads are scarce
        ctx.set_option("raw_inputs", 2)
        for plan, threads, expect_raw in ((1, 1, None), (1, 16, None), (0, 2, True), (0, 16, False)):
            ctx.set_option("host_threads", threads); ctx.set_option("device_plan", plan)
            ctx.reset_stats()
            r7, c7 = ctx.sw_extend_batch(p, *flat)
            assert_same(ro, r7, f"auto, {threads} host threads, device_plan {plan}")
            assert_same(co.astype(np.int64), c7.astype(np.int64), "cells")
            if expect_raw is not None:
                assert (ctx.stats()["h2d_bytes"] > int(t["qoff"][-1] + t["toff"][-1])) == expect_raw
        ctx.set_option("device_plan", 1)
        # 2 bit per base forced: N tasks are rerun, a bad code is reported with its task
        ctx.set_option("raw_inputs", 3)
        ctx.reset_stats()
        r9, c9 = ctx.sw_extend_batch(p, *flat)
        assert_same(ro, r9, "2-bit lean path"); assert_same(co.astype(np.int64), c9.astype(np.int64), "2-bit lean path cells")
        assert ctx.stats()["h2d_bytes"] < (int(t["qoff"][-1] + t["toff"][-1])) // 2 + 40 * len(ro)
        keep = t["qbuf"][int(t["qoff"][777]) + 1]
        t["qbuf"][int(t["qoff"][777]) + 1] = 7
        with pytest.raises(B.BswError) as e2:
            ctx.sw_extend_batch(p, *flat)
        t["qbuf"][int(t["qoff"][777]) + 1] = keep
        assert e2.value.code == B.BSW_EINVAL and "task 777" in str(e2.value)
        ctx.set_option("raw_inputs", 2)
        # the caller's result array page-locked as well: the records are copied straight into it
        out = np.zeros(len(ro), dtype=B.RESULT_DTYPE)
        ctx.register_host(out)
        try:
            r8, c8 = ctx.sw_extend_batch(p, *flat, out=out)
            assert r8 is out or np.shares_memory(r8, out)
            assert_same(ro, out, "results DMA-ed into the registered array"); assert_same(co.astype(np.int64), c8.astype(np.int64), "cells")
        finally:
            ctx.unregister_host(out)
        ctx.set_option("host_threads", 0)
    finally:
        ctx.set_option("raw_inputs", 2); ctx.set_option("host_threads", 0)
        ctx.unregister_host(t["qbuf"]); ctx.unregister_host(t["tbuf"])


def test_level3_wire_format(B, O, ctx):
    """TBB image in, RBB image out (the FPGA has no z-drop and a fixed +1/-4/-1 matrix: compare with zdrop=0)."""
    t = B.synth_tasks("cfg1_101bp", 1600, seed=30)
    seeds = seeds_from_flat(t, 800, unset_score_every=4)
    P2 = B.make_params2(B.make_params(zdrop=0), w=100, pen_clip5=5, pen_clip3=5)
    want, _ = oracle_chain2aln(O, B, P2, seeds)
    tbb = B.tbb_encode(P2, seeds)
    rbb, n = ctx.pe_array_batch(tbb)
    assert n == 800
    got = B.rbb_decode(rbb, n)
    assert_same(want, got, "RBB records")
    assert not rbb[5 * n:].any()                   # words past the records are left untouched


def test_full_size_properties(B, O, ctx):
    """BASELINE configs[1] at full size (1M x 150 bp): size-independent properties + a 1% oracle sample."""
    n = 1_000_000
    t = B.synth_tasks("cfg2_150bp", n, seed=1)
    p, po = B.make_params(), O.make_params()
    ctx.reset_stats()
    r, c = ctx.sw_extend_batch(p, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    st = ctx.stats()
    qlen = np.diff(t["qoff"]); tlen = np.diff(t["toff"])
    assert st["tasks"] == n and st["cells_band"] == int(c.astype(np.int64).sum())       # device counter == sum of per-task cells
    assert (r["score"] >= t["h0"]).all() and (r["score"] <= t["h0"] + qlen).all()
    assert (r["qle"] >= 0).all() and (r["qle"] <= qlen).all() and (r["tle"] >= 0).all() and (r["tle"] <= tlen).all()
    assert (r["gtle"] <= tlen).all() and (r["gscore"] <= r["score"]).all() and (r["max_off"] >= 0).all()
    assert (c <= qlen.astype(np.int64) * tlen).all()
    # idempotence + permutation invariance (the scheduler reorders tasks internally)
    perm = np.random.default_rng(0).permutation(n)[:50_000]
    qs = np.concatenate([[0], np.cumsum(qlen[perm])]); ts = np.concatenate([[0], np.cumsum(tlen[perm])])
    qb = np.concatenate([t["qbuf"][t["qoff"][i]:t["qoff"][i + 1]] for i in perm] + [np.zeros(8, np.uint8)])
    tb = np.concatenate([t["tbuf"][t["toff"][i]:t["toff"][i + 1]] for i in perm] + [np.zeros(8, np.uint8)])
    r2, c2 = ctx.sw_extend_batch(p, qb, qs, tb, ts, t["h0"][perm], t["w"][perm])
    assert_same(r[perm], r2, "permutation invariance")
    assert_same(c[perm], c2, "cells under permutation")
    # oracle on the permuted 5% sample
    ro, co = O.extend_batch(po, qb, qs, tb, ts, t["h0"][perm], t["w"][perm])
    assert_same(ro, r2, "oracle sample")
    assert_same(co, c2.astype(np.int64), "oracle sample cells")


def test_int_peak_microbenchmark(ctx):
    pk = ctx.measure_int_peak(0)
    assert pk["sm_count"] >= 100 and 5.0 < pk["iadd_tops"] < 80.0 and pk["dpx_tops"] > pk["vimnmx_tops"]


def test_multi_device_context_shards_without_collective(B, O):
    """One context over every visible GPU: chunks are pulled dynamically by workers bound round-robin to the devices."""
    import torch
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    t = B.synth_tasks("cfg3_mixed", 200_000, seed=40)
    p, po = B.make_params(), O.make_params()
    with B.Context(devices=list(range(ndev)), chunk_tasks=8192) as mctx:
        assert mctx.num_devices == ndev
        rg, cg = mctx.sw_extend_batch(p, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    ro, co = O.extend_batch(po, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    assert_same(ro, rg, "multi-device results")
    assert_same(co, cg.astype(np.int64), "multi-device cells")


def test_level2_host_orchestrated_and_long_flanks(B, O, ctx):
    """fused_l2=0 (four level-1 passes) gives the same records; seeds with a flank beyond a K1 tile take that path anyway."""
    t = B.synth_tasks("cfg3_mixed", 2000, seed=70)
    seeds = seeds_from_flat(t, 1000, unset_score_every=3)
    lt = B.synth_tasks("cfg4_long", 8, seed=71)                    # 4 seeds with 1-10 kb flanks
    for r in range(4):
        l, g = 2 * r, 2 * r + 1
        seeds.append(dict(q_left=lt["qbuf"][lt["qoff"][l]:lt["qoff"][l + 1]], t_left=lt["tbuf"][lt["toff"][l]:lt["toff"][l + 1]],
                          q_right=lt["qbuf"][lt["qoff"][g]:lt["qoff"][g + 1]], t_right=lt["tbuf"][lt["toff"][g]:lt["toff"][g + 1]],
                          init_score=-1, qbeg=int(lt["qoff"][l + 1] - lt["qoff"][l]), h0=int(lt["h0"][l]), id=9000 + r))
    P2 = B.make_params2(B.make_params(zdrop=100), w=100, pen_clip5=5, pen_clip3=5)
    want, _ = oracle_chain2aln(O, B, P2, seeds)
    got = ctx.proc_element_batch(P2, seeds)
    assert_same(want, got, "fused + leftover")
    ctx.set_option("fused_l2", 0)
    try:
        got2 = ctx.proc_element_batch(P2, seeds)
    finally:
        ctx.set_option("fused_l2", 1)
    assert_same(want, got2, "host-orchestrated")


def test_level2_seeds_beyond_16_bits(B, O, ctx):
    """Seed tasks whose chain already scores more than the 16-bit row state holds (a 60 kb read: init_score = seed
    length x a): they leave the fused kernel for the host-orchestrated path, whose extension calls put them on K5."""
    t = B.synth_tasks("cfg3_mixed", 2000, seed=72)
    seeds = seeds_from_flat(t, 1000, unset_score_every=3)
    for r in (3, 4, 5, 6, 500, 999):
        s = dict(seeds[r]); s["h0"] = 60_000 + r
        if s["init_score"] >= 0: s["init_score"] = 60_000 + r
        seeds[r] = s
    P2 = B.make_params2(B.make_params(zdrop=100), w=100, pen_clip5=5, pen_clip3=5)
    want, _ = oracle_chain2aln(O, B, P2, seeds)
    got = ctx.proc_element_batch(P2, seeds)
    assert_same(want, got, "fused + 32-bit leftovers")
    assert int(want["score"].max()) > 60_000


def test_long_and_wide_rows_on_k2(B, O, ctx):
    """K2 on a mix that crosses its thresholds (k2_min_qlen = 64), other gap penalties with N bases, near-perfect 3 kb
    matches whose windows grow to the band (one and four warps, with and without the narrow-row path)."""
    both(B, O, ctx, B.synth_tasks("cfg3_mixed", 20_000, seed=80), opts={"k2_min_qlen": 64})
    both(B, O, ctx, B.synth_tasks("cfg3_mixed", 5_000, seed=81, n_frac=0.02), opts={"k2_min_qlen": 64}, o_del=4, e_del=2, o_ins=7, e_ins=1)
    rng = np.random.default_rng(5)
    qs, ts, h0, w = [], [], [], []
    for k in range(40):                                       # near-perfect 3 kb matches with a big h0: windows beyond 510 columns
        q = rng.integers(0, 4, 3000).astype(np.uint8)
        t = np.concatenate([q, q[:200]]).astype(np.uint8)
        t[rng.random(len(t)) < 0.01] = 0
        qs.append(q); ts.append(t); h0.append(400 if k % 2 == 0 else 30); w.append(400)
    qbuf, qoff, tbuf, toff = flat_from_lists(qs, ts)
    t = dict(qbuf=qbuf, qoff=qoff, tbuf=tbuf, toff=toff, h0=np.array(h0, np.int32), w=np.array(w, np.int32))
    both(B, O, ctx, t)
    both(B, O, ctx, B.synth_tasks("cfg4_long", 16, seed=84))
    both(B, O, ctx, B.synth_tasks("cfg4_long", 32, seed=82))          # default: K2 (one warp per task, ring row buffer)
    both(B, O, ctx, B.synth_tasks("cfg4_long", 16, seed=83), opts={"k2_warps": 4})
    ctx.set_option("k2_warps", 1)
    rng = np.random.default_rng(6)
    qs, ts = [], []
    for k in range(12):                                       # near-perfect 3 kb matches: rows as wide as the band allows
        q = rng.integers(0, 4, 3000).astype(np.uint8)
        qs.append(q); ts.append(np.concatenate([q, q[:200]]).astype(np.uint8))
    qbuf, qoff, tbuf, toff = flat_from_lists(qs, ts)
    wide = dict(qbuf=qbuf, qoff=qoff, tbuf=tbuf, toff=toff, h0=np.full(12, 400, np.int32), w=np.full(12, 400, np.int32))
    both(B, O, ctx, wide)
    both(B, O, ctx, wide, opts={"k2_narrow": 0}); ctx.set_option("k2_narrow", 1)


@pytest.mark.parametrize("seed", [1, 2])
def test_randomised_small_tasks_every_kernel(B, O, ctx, seed):
    """Adversarial little tasks (h0 down to 1, w down to 0, ties, indels), random scoring, through K1, K2 (1 and 4
    warps, both variants, with and without the narrow-row path) and the fused level 2."""
    from helpers import random_small_tasks
    rng = np.random.default_rng(2000 + seed)
    t = random_small_tasks(rng, 4000)
    pk = dict(o_del=int(rng.integers(0, 6)), e_del=int(rng.integers(1, 4)), o_ins=int(rng.integers(0, 6)), e_ins=int(rng.integers(1, 4)),
              zdrop=int(rng.choice([0, 3, 10, 100])), a=int(rng.integers(1, 4)), b=int(rng.integers(1, 6)), end_bonus=int(rng.integers(0, 8)))
    for p in (dict(), pk):
        both(B, O, ctx, t, **p)
        both(B, O, ctx, t, variant=2, **p)
        both(B, O, ctx, t, opts={"k2_min_qlen": 8}, **p)
        both(B, O, ctx, t, opts={"force_kernel": 2}, **p)
        both(B, O, ctx, t, opts={"force_kernel": 2, "k2_warps": 4}, **p); ctx.set_option("k2_warps", 1)
        both(B, O, ctx, t, opts={"force_kernel": 2, "k2_narrow": 0}, **p); ctx.set_option("k2_narrow", 1)
        both(B, O, ctx, t, variant=2, opts={"force_kernel": 2}, **p)
    # level 2 on random flank pairs
    n = 2000
    seeds = []
    for r in range(n // 2):
        l, g = 2 * r, 2 * r + 1
        sl = lambda a, off, i: a[off[i]:off[i + 1]]
        ql, tl, qr, tr = sl(t["qbuf"], t["qoff"], l), sl(t["tbuf"], t["toff"], l), sl(t["qbuf"], t["qoff"], g), sl(t["tbuf"], t["toff"], g)
        if r % 5 == 1:
            ql, tl = ql[:0], tl[:0]
        if r % 5 == 2:
            qr, tr = qr[:0], tr[:0]
        h0 = int(t["h0"][l])
        seeds.append(dict(q_left=ql, q_right=qr, t_left=tl, t_right=tr, init_score=(h0 if len(ql) == 0 else -1), qbeg=len(ql), h0=h0, id=r))
    for w in (3, 100):
        P2 = B.make_params2(B.make_params(**pk), w=w, pen_clip5=int(rng.integers(0, 8)), pen_clip3=int(rng.integers(0, 8)))
        want, _ = oracle_chain2aln(O, B, P2, seeds)
        assert_same(want, ctx.proc_element_batch(P2, seeds), "fused level 2")


# ---------------------------------------------------------------- BASELINE configs at their full size, every task compared
@pytest.mark.parametrize("name,n", [("cfg2_150bp", 1_000_000), ("cfg3_mixed", 1_000_000)])
def test_full_size_every_task_every_field(B, O, ctx, name, n):
    """BASELINE configs[1] / [2] at full size: all six outputs and the per-task cell count of ALL tasks against the oracle."""
    both(B, O, ctx, B.synth_tasks(name, n, seed=1))


@pytest.mark.parametrize("zdrop", [100, 400])
def test_cfg4_all_long_reads(B, O, ctx, zdrop):
    """BASELINE configs[3]: all 20 000 long-read extensions (1-10 kb, w = 500) on the intra-task kernel."""
    both(B, O, ctx, B.synth_tasks("cfg4_long", 20_000, seed=1), zdrop=zdrop)


@pytest.mark.parametrize("variant", [1, 2])
def test_scores_at_the_16_bit_cap(B, O, ctx, variant):
    """h0 = 32767 - qlen*max(mat): the packed 16x2 cell adds (a+b) before it subtracts b, i.e. it wraps inside the
    intrinsic when M is within a+b of the cap; the admission rule promises exactness up to 32767."""
    rng = np.random.default_rng(77)
    qs, ts, h0 = [], [], []
    for k in range(3000):
        ql = int(rng.integers(1, 200))
        q = rng.integers(0, 4, ql).astype(np.uint8)
        t = np.concatenate([q, rng.integers(0, 4, int(rng.integers(0, 30))).astype(np.uint8)])    # perfect match: the score climbs to the cap
        if k % 3 == 1:
            t[rng.integers(0, len(t), max(1, len(t) // 20))] = rng.integers(0, 4)
        if k % 7 == 3:
            q[int(rng.integers(0, ql))] = 4                                                           # N: matrix-lookup kernel
        qs.append(q); ts.append(t); h0.append(32767 - ql * 1)
    qbuf, qoff, tbuf, toff = flat_from_lists(qs, ts)
    t = dict(qbuf=qbuf, qoff=qoff, tbuf=tbuf, toff=toff, h0=np.array(h0, np.int32), w=np.full(len(h0), 100, np.int32))
    ro, _ = both(B, O, ctx, t, variant=variant, zdrop=0)
    assert int(ro["score"].max()) == 32767
    both(B, O, ctx, t, variant=variant, opts={"force_kernel": 2}, zdrop=0)                            # K2's packed cell wraps the same way
    # a=2: max(mat) = 2
    t["h0"] = np.array([32767 - 2 * len(q) for q in qs], np.int32)
    ro, _ = both(B, O, ctx, t, variant=variant, a=2, b=3, zdrop=0)
    assert int(ro["score"].max()) == 32767
    both(B, O, ctx, t, variant=variant, opts={"force_kernel": 2}, a=2, b=3, zdrop=0)
    t["h0"][5] += 1                                                                                   # one past the cap: never wrapped --
    ctx.set_option("wide", 0)                                                                         # refused without the 32-bit kernel,
    try:
        with pytest.raises(B.BswError) as e:
            ctx.sw_extend_batch(B.make_params(a=2, b=3, zdrop=0), t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
        assert e.value.code == B.BSW_ERANGE
    finally:
        ctx.set_option("wide", 1)
    ro, _ = both(B, O, ctx, t, variant=variant, a=2, b=3, zdrop=0)                                    # exact with it (that one task runs on K5)
    assert int(ro["score"].max()) == 32768


@pytest.mark.parametrize("variant", [1, 2])
def test_long_reads_at_the_16_bit_cap(B, O, ctx, variant):
    """K2 in its own regime at the edge of the envelope: 3-8 kb near-perfect matches whose score climbs to exactly 32767
    (h0 = 32767 - qlen), windows 250-400 columns wide (the 12-column rounds), ring row buffer."""
    rng = np.random.default_rng(78)
    qs, ts, h0 = [], [], []
    for k in range(24):
        ql = int(rng.integers(3000, 8000))
        q = rng.integers(0, 4, ql).astype(np.uint8)
        t = np.concatenate([q, rng.integers(0, 4, 60).astype(np.uint8)])
        if k % 2:
            t[rng.integers(0, ql, ql // 200)] = rng.integers(0, 4)                # 0.5 % substitutions
        qs.append(q); ts.append(t); h0.append(32767 - ql)
    qbuf, qoff, tbuf, toff = flat_from_lists(qs, ts)
    t = dict(qbuf=qbuf, qoff=qoff, tbuf=tbuf, toff=toff, h0=np.array(h0, np.int32), w=np.full(len(h0), 180, np.int32))
    ro, _ = both(B, O, ctx, t, variant=variant, zdrop=0)
    assert int(ro["score"].max()) == 32767
    both(B, O, ctx, t, variant=variant, zdrop=100)


@pytest.mark.parametrize("variant", [1, 2])
def test_every_task_on_the_32_bit_kernel(B, O, ctx, variant):
    """Option wide=2 sends every task to K5 (32-bit rows in global memory, one warp per task): the usual corpora must
    come back bit-exact, cells included -- short reads, N bases, other scorings, z-drop on and off, long reads."""
    from helpers import random_small_tasks
    try:
        both(B, O, ctx, B.synth_tasks("cfg3_mixed", 8_000, seed=31, n_frac=0.01), variant=variant, opts={"wide": 2})
        both(B, O, ctx, B.synth_tasks("cfg2_150bp", 8_000, seed=32), variant=variant, opts={"wide": 2}, zdrop=0)
        both(B, O, ctx, random_small_tasks(np.random.default_rng(33), 6000), variant=variant, opts={"wide": 2})
        both(B, O, ctx, B.synth_tasks("cfg3_mixed", 4_000, seed=34), variant=variant, opts={"wide": 2}, a=2, b=3, o_del=4, e_del=2, o_ins=7, e_ins=1, zdrop=20)
        both(B, O, ctx, B.synth_tasks("cfg4_long", 200, seed=35), variant=variant, opts={"wide": 2})
    finally:
        ctx.set_option("wide", 1)


@pytest.mark.parametrize("variant", [1, 2])
def test_tasks_beyond_16_bits_in_a_mixed_batch(B, O, ctx, variant):
    """A batch of ordinary tasks with a few that only 32-bit rows can hold: h0 far above 32767 (scores above 16 bits,
    with and without mismatches and gaps) and a 45 kb query (longer than K2's shared-memory row).  The call splits the
    batch -- K1 / K2 for the ordinary tasks, K5 for the rest -- and every task equals the oracle."""
    rng = np.random.default_rng(91)
    t = B.synth_tasks("cfg3_mixed", 3000, seed=36)
    qs = [t["qbuf"][t["qoff"][i]:t["qoff"][i + 1]] for i in range(3000)]
    ts = [t["tbuf"][t["toff"][i]:t["toff"][i + 1]] for i in range(3000)]
    h0 = list(t["h0"]); w = list(t["w"])
    for k, ql in enumerate((200, 3000, 45_000, 700, 12_000)):
        q = rng.integers(0, 4, ql).astype(np.uint8)
        tt = q.copy()
        tt[rng.integers(0, ql, max(1, ql // 50))] = rng.integers(0, 4)              # 2 % substitutions
        cut = int(rng.integers(ql // 3, ql // 2))
        tt = np.concatenate([tt[:cut], rng.integers(0, 4, 7).astype(np.uint8), tt[cut:], rng.integers(0, 4, 40).astype(np.uint8)])
        pos = 17 * k + 5
        qs.insert(pos, q); ts.insert(pos, tt)
        h0.insert(pos, [70_000, 33_000, 19, 1_000_000, 40_000][k]); w.insert(pos, [100, 150, 200, 50, 300][k])
    qbuf, qoff, tbuf, toff = flat_from_lists(qs, ts)
    tt = dict(qbuf=qbuf, qoff=qoff, tbuf=tbuf, toff=toff, h0=np.array(h0, np.int32), w=np.array(w, np.int32))
    ro, _ = both(B, O, ctx, tt, variant=variant)
    assert int(ro["score"].max()) > 1_000_000
    ro, _ = both(B, O, ctx, tt, variant=variant, zdrop=0)
