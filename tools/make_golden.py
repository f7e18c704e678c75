"""Generate tests/golden/*.npz: small frozen input/output vectors for the level-1 and level-2 calls.

The reference tree holds no golden vectors, and its RTL cannot be simulated here, so these
fixtures are produced by the CPU oracle (oracle/ksw_extend_ref.c) after it has passed tests/test_oracle_kat.py; they
freeze its behaviour so that a later edit of the oracle, the scheduler or the kernels is caught.  Re-run:
    python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bsw_b200 as B  # noqa: E402
import oracle as O  # noqa: E402
from helpers import oracle_chain2aln, seeds_from_flat  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def l1(name, workload, n, seed, variant, n_frac=0.0, **pk):
    t = B.synth_tasks(workload, n, seed=seed, n_frac=n_frac)
    po = O.make_params(**pk)
    res, cells = O.extend_batch(po, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"], variant=variant)
    mat = np.array([po.mat[i] for i in range(25)], dtype=np.int8)
    scal = np.array([po.o_del, po.e_del, po.o_ins, po.e_ins, po.zdrop, po.end_bonus, variant], dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, name), qbuf=t["qbuf"], qoff=t["qoff"], tbuf=t["tbuf"], toff=t["toff"], h0=t["h0"],
                        w=t["w"], mat=mat, scal=scal, res=res, cells=cells)
    print(name, n, "tasks", int(cells.sum()), "cells")


def l2(name, workload, nreads, seed):
    t = B.synth_tasks(workload, 2 * nreads, seed=seed)
    seeds = seeds_from_flat(t, nreads, unset_score_every=3)
    P2 = B.make_params2(w=100, pen_clip5=5, pen_clip3=5)
    out, cells = oracle_chain2aln(O, B, P2, seeds)
    np.savez_compressed(os.path.join(OUT, name), qbuf=t["qbuf"], qoff=t["qoff"], tbuf=t["tbuf"], toff=t["toff"], h0=t["h0"],
                        nreads=np.int64(nreads), rec=out, cells=np.int64(cells))
    print(name, nreads, "seed tasks", cells, "cells")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    l1("l1_cfg1_v1.npz", "cfg1_101bp", 300, 11, 1)
    l1("l1_cfg3_v1.npz", "cfg3_mixed", 300, 12, 1)
    l1("l1_cfg3_v2.npz", "cfg3_mixed", 300, 12, 2)
    l1("l1_cfg3_n_asym_v1.npz", "cfg3_mixed", 200, 13, 1, n_frac=0.02, o_del=4, e_del=2, o_ins=7, e_ins=1, zdrop=30)
    l1("l1_cfg4_long_v1.npz", "cfg4_long", 4, 14, 1)
    l2("l2_cfg3.npz", "cfg3_mixed", 150, 15)
