"""Frozen vectors (tests/golden, made by tools/make_golden.py): the oracle and the CPU emulation must reproduce them
here; the CUDA path must reproduce them on the GPU (test_gpu_parity.py::test_golden_*)."""
import glob
import os

import numpy as np
import pytest

from helpers import assert_same, oracle_chain2aln, seeds_from_flat

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
L1 = sorted(glob.glob(os.path.join(GOLD, "l1_*.npz")))


def params_of(mod, g):
    s = g["scal"]
    return mod.make_params(mat=g["mat"], o_del=int(s[0]), e_del=int(s[1]), o_ins=int(s[2]), e_ins=int(s[3]),
                           zdrop=int(s[4]), end_bonus=int(s[5])), int(s[6])


def test_fixtures_exist():
    assert len(L1) >= 5 and os.path.exists(os.path.join(GOLD, "l2_cfg3.npz"))


@pytest.mark.parametrize("path", L1, ids=[os.path.basename(p) for p in L1])
def test_oracle_reproduces_golden(O, path):
    g = np.load(path)
    po, variant = params_of(O, g)
    res, cells = O.extend_batch(po, g["qbuf"], g["qoff"], g["tbuf"], g["toff"], g["h0"], g["w"], variant=variant)
    assert_same(res, g["res"], "oracle vs golden")
    assert_same(cells, g["cells"], "cells")


@pytest.mark.parametrize("path", [p for p in L1 if "long" not in p], ids=[os.path.basename(p) for p in L1 if "long" not in p])
def test_emulation_reproduces_golden(B, path):
    g = np.load(path)
    p, variant = params_of(B, g)
    res, cells, _ = B.emu_extend_batch(p, g["qbuf"], g["qoff"], g["tbuf"], g["toff"], g["h0"], g["w"], variant=variant)
    assert_same(res, g["res"], "emulation vs golden")
    assert_same(cells.astype(np.int64), g["cells"], "cells")


def test_oracle_reproduces_level2_golden(O, B):
    g = np.load(os.path.join(GOLD, "l2_cfg3.npz"))
    seeds = seeds_from_flat(g, int(g["nreads"]), unset_score_every=3)
    out, cells = oracle_chain2aln(O, B, B.make_params2(w=100, pen_clip5=5, pen_clip3=5), seeds)
    assert_same(out, g["rec"], "level 2")
    assert cells == int(g["cells"])
