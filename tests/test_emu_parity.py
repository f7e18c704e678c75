"""CPU-side checks of the product's host logic (scheduler, packer) and of the K1 lane function's control flow, by
running csrc/emu.cpp (the same bsw_k1_core.cuh the device kernel instantiates, compiled for the host) against the
oracle.  The CUDA kernels themselves are checked in test_gpu_parity.py."""
import numpy as np
import pytest

from helpers import assert_same, flat_from_lists


def run_both(B, O, t, variant=1, **pk):
    p, po = B.make_params(**pk), O.make_params(**pk)
    ro, co = O.extend_batch(po, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"], variant=variant)
    re, ce, info = B.emu_extend_batch(p, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"], variant=variant)
    assert_same(ro, re, "results")
    assert_same(co.astype(np.int64), ce.astype(np.int64), "cells")
    return info


@pytest.mark.parametrize("name", ["cfg1_101bp", "cfg2_150bp", "cfg3_mixed"])
@pytest.mark.parametrize("variant", [1, 2])
def test_workloads(B, O, name, variant):
    info = run_both(B, O, B.synth_tasks(name, 6000, seed=variant), variant=variant)
    assert info[1] == (6000 + 31) // 32          # one tile per 32 tasks, no N -> single class


@pytest.mark.parametrize("variant", [1, 2])
def test_ambiguous_bases_use_matrix_lookup(B, O, variant):
    info = run_both(B, O, B.synth_tasks("cfg3_mixed", 4000, n_frac=0.02), variant=variant)
    assert info[3] > 0 or info[1] > 125          # two classes (N-free / with N) -> padded lanes or an extra tile


@pytest.mark.parametrize("pk", [dict(o_del=4, e_del=2, o_ins=7, e_ins=1), dict(a=2, b=3, zdrop=20),
                                dict(zdrop=0, o_del=0, o_ins=0), dict(o_del=0, e_del=3, o_ins=9, e_ins=2, end_bonus=0)])
@pytest.mark.parametrize("variant", [1, 2])
def test_scoring_variations(B, O, pk, variant):
    run_both(B, O, B.synth_tasks("cfg3_mixed", 3000, seed=3), variant=variant, **pk)


def test_custom_matrix(B, O):
    m = B.bwa_fill_scmat(1, 4).copy()
    m[1], m[7], m[24] = 2, -3, 0
    run_both(B, O, B.synth_tasks("cfg3_mixed", 3000, seed=5), mat=m)


def test_edge_shapes(B, O):
    rng = np.random.default_rng(9)
    qs, ts, h0, w = [], [], [], []
    for qlen, tlen, h, ww in [(1, 1, 1, 100), (1, 50, 30, 100), (50, 1, 30, 100), (8, 8, 5, 0), (9, 200, 100, 3),
                              (16, 16, 1, 100), (17, 40, 300, 1), (255, 600, 19, 100), (300, 310, 1000, 50)]:
        q = rng.integers(0, 4, qlen).astype(np.uint8)
        t = np.resize(q, tlen).astype(np.uint8)
        t[rng.random(tlen) < 0.1] = 3
        qs.append(q); ts.append(t); h0.append(h); w.append(ww)
    qbuf, qoff, tbuf, toff = flat_from_lists(qs, ts)
    t = dict(qbuf=qbuf, qoff=qoff, tbuf=tbuf, toff=toff, h0=np.array(h0, np.int32), w=np.array(w, np.int32))
    for variant in (1, 2):
        run_both(B, O, t, variant=variant)


def test_rejects_bad_input(B):
    p = B.make_params()
    q = np.array([0, 1, 7, 2], np.uint8)                      # base code 7
    qbuf, qoff, tbuf, toff = flat_from_lists([q], [q])
    with pytest.raises(B.BswError) as e:
        B.emu_extend_batch(p, qbuf, qoff, tbuf, toff, [10], [100])
    assert e.value.code == B.BSW_EINVAL
    q = np.zeros(40000, np.uint8)                             # h0 + qlen*a exceeds the 16-bit row state
    qbuf, qoff, tbuf, toff = flat_from_lists([q], [q[:10]])
    with pytest.raises(B.BswError) as e:
        B.emu_extend_batch(p, qbuf, qoff, tbuf, toff, [10], [100])
    assert e.value.code == B.BSW_ERANGE


@pytest.mark.parametrize("w,zdrop,variant", [(100, 100, 1), (10, 100, 1), (5, 0, 1), (100, 100, 2), (7, 50, 2)])
def test_fused_seed_task_lane_function(B, O, w, zdrop, variant):
    """K3 (bsw_k3_core.cuh): left + right extension, band retry, clip decision per lane == the oracle's chain2aln."""
    from helpers import oracle_chain2aln, seeds_from_flat
    t = B.synth_tasks("cfg3_mixed", 3000, seed=60 + w, n_frac=0.005)
    seeds = seeds_from_flat(t, 1500, unset_score_every=3)
    P2 = B.make_params2(B.make_params(zdrop=zdrop), w=w, pen_clip5=5, pen_clip3=7)
    want, _ = oracle_chain2aln(O, B, P2, seeds, variant=variant)
    got = B.emu_chain2aln(P2, seeds, variant=variant)
    assert_same(want, got, "fused seed task")
    if w < 100:
        assert (want["w"] == 2 * w).any()          # the band-retry path ran


def _emu_flags(B, **kw):
    import ctypes as C
    return {k: C.c_int.in_dll(B.emu_lib(), "bsw_emu_" + k) for k in kw}


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_randomised_small_tasks(B, O, seed):
    """Thousands of adversarial little tasks (h0 down to 1, w down to 0, ties, indels) under random scoring, both variants,
    through the K1 lane function."""
    from helpers import random_small_tasks
    rng = np.random.default_rng(1000 + seed)
    t = random_small_tasks(rng, 3000)
    pks = [dict(), dict(o_del=int(rng.integers(0, 6)), e_del=int(rng.integers(1, 4)), o_ins=int(rng.integers(0, 6)), e_ins=int(rng.integers(1, 4)),
                     zdrop=int(rng.choice([0, 3, 10, 100])), a=int(rng.integers(1, 4)), b=int(rng.integers(1, 6)), end_bonus=int(rng.integers(0, 8)))]
    for pk in pks:
        for variant in (1, 2):
            run_both(B, O, t, variant=variant, **pk)
