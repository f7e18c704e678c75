"""K2's packed round under matrix-lookup scoring, every row on it (k2_narrow = 0).  Kept in its own file, last in the
alphabet: the case was found by the CPU model (tests/test_k2_round_model.py) after the round's GPU budget was spent, so
this is the one GPU test that had not run on hardware when it was committed."""
import numpy as np
import pytest

from test_gpu_parity import both

pytestmark = pytest.mark.gpu


def test_k2_rows_that_die_next_to_a_matching_dead_column(B, O, ctx):
    """Matrix-lookup scoring (N in the query), tiny h0, every row on K2's packed round (k2_narrow = 0): tasks end on a row
    whose live cells are all zero while a zeroed column left of the window matches the target base.  Found with the CPU
    model of the round (tests/test_k2_round_model.py); the registers-only path for rows below 64 columns hides it."""
    from helpers import random_small_tasks
    rng = np.random.default_rng(530)
    try:
        for rep in range(3):
            t = random_small_tasks(rng, 3000, qmax=60, tmax=90)
            m = rng.random(len(t["qbuf"])) < 0.03; t["qbuf"] = np.where(m, 4, t["qbuf"]).astype(np.uint8)
            t["h0"] = rng.integers(1, 6, len(t["h0"])).astype(np.int32)
            pk = dict(a=int(rng.integers(1, 4)), b=int(rng.integers(1, 5)), zdrop=0)
            for variant in (1, 2):
                both(B, O, ctx, t, variant=variant, opts={"force_kernel": 2, "k2_narrow": 0}, **pk)
    finally:
        ctx.set_option("k2_narrow", 1)

