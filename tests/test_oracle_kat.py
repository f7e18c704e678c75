"""The oracle (oracle/ksw_extend_ref.c) against hand-derivable known answers and an independently structured
full-matrix model (oracle/matrix_model.py).  The reference tree holds no tests or vectors for this path
(SURVEY.md section 4), so these KATs are what pins the restatement ("parity unpinned" by the reference itself)."""
import numpy as np
import pytest

from oracle import matrix_model as MM


def ext(O, q, t, h0, w=100, variant=1, **kw):
    r, cells = O.extend_one(O.make_params(**kw), q, t, h0, w, variant)
    return {k: int(r[k]) for k in r.dtype.names}, cells


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("qlen,h0", [(1, 1), (10, 20), (40, 19), (101, 60)])
def test_perfect_match(O, variant, qlen, h0):
    rng = np.random.default_rng(qlen)
    q = rng.integers(0, 4, qlen).astype(np.uint8)
    r, _ = ext(O, q, q, h0, variant=variant)
    assert r == dict(score=h0 + qlen, qle=qlen, tle=qlen, gtle=qlen, gscore=h0 + qlen, max_off=0)


@pytest.mark.parametrize("variant", [1, 2])
def test_all_mismatch(O, variant):
    q = np.zeros(30, np.uint8)          # AAAA...
    t = np.ones(30, np.uint8)           # CCCC...
    r, _ = ext(O, q, t, 25, variant=variant)
    assert r["score"] == 25 and r["qle"] == 0 and r["tle"] == 0 and r["max_off"] == 0


@pytest.mark.parametrize("variant", [1, 2])
def test_single_deletion_and_insertion(O, variant):
    rng = np.random.default_rng(7)
    q = rng.integers(0, 4, 60).astype(np.uint8)
    # one extra target base after 30 matched bases: 60 matches, one gap of length 1 (o=6, e=1) -> h0 + 60 - 7
    extra = np.uint8((q[30] + 1) % 4 if (q[30] + 1) % 4 != q[29] else (q[30] + 2) % 4)
    t = np.concatenate([q[:30], [extra], q[30:]]).astype(np.uint8)
    r, _ = ext(O, q, t, 40, variant=variant)
    assert r["score"] == 40 + 60 - 7 and r["qle"] == 60 and r["tle"] == 61 and r["max_off"] == 1
    assert r["gscore"] == r["score"] and r["gtle"] == 61
    # one query base missing from the target
    t2 = np.concatenate([q[:30], q[31:]]).astype(np.uint8)
    r2, _ = ext(O, q, t2, 40, variant=variant)
    assert r2["score"] == 40 + 59 - 7 and r2["qle"] == 60 and r2["tle"] == 59 and r2["max_off"] == 1


def test_n_bases_score_minus_one(O):
    q = np.array([0, 1, 2, 3, 4, 0, 1, 2, 3, 0], np.uint8)     # one N in the query
    r, _ = ext(O, q, q.copy(), 20)
    # N vs N scores -1 (the N row/column of the matrix): 9 matches - 1
    assert r["score"] == 20 + 9 - 1 and r["qle"] == 10 and r["tle"] == 10


def test_target_shorter_than_query(O):
    rng = np.random.default_rng(3)
    q = rng.integers(0, 4, 50).astype(np.uint8)
    r, _ = ext(O, q, q[:20], 30)
    assert r["score"] == 50 and r["qle"] == 20 and r["tle"] == 20
    assert r["gscore"] <= 0 or r["gtle"] <= 20


def test_zdrop_stops_early(O):
    rng = np.random.default_rng(5)
    q = rng.integers(0, 4, 200).astype(np.uint8)
    t = np.concatenate([q[:50], rng.integers(0, 4, 250)]).astype(np.uint8)
    r_on, c_on = ext(O, q, t, 60, zdrop=10)
    r_off, c_off = ext(O, q, t, 60, zdrop=0)
    assert r_on["score"] == r_off["score"] and 110 <= r_on["score"] <= 114      # 50 matches + a few chance matches
    assert c_on < c_off


def test_band_clamp_formula(O):
    p = O.make_params()
    # max_ins = (qlen*1 + 5 - 6)/1 + 1 = qlen  -> w = min(100, qlen)
    assert O.lib().bswref_clamp_w(p, 30, 100, 5) == 30
    assert O.lib().bswref_clamp_w(p, 300, 100, 5) == 100
    assert O.lib().bswref_clamp_w(p, 1, 100, 5) == 1


@pytest.mark.parametrize("variant", [1, 2])
def test_against_matrix_model(O, variant):
    """Row-buffer oracle == full-matrix model on random small tasks (related, diverged and unrelated targets)."""
    rng = np.random.default_rng(100 + variant)
    mat = O.bwa_fill_scmat(1, 4)
    for trial in range(250):
        qlen = int(rng.integers(1, 45))
        tlen = int(rng.integers(1, 70))
        q = rng.integers(0, 4, qlen).astype(np.uint8)
        mode = trial % 3
        if mode == 0:
            t = rng.integers(0, 4, tlen).astype(np.uint8)
        else:
            t = np.resize(q, tlen).astype(np.uint8)
            flips = rng.random(tlen) < (0.05 if mode == 1 else 0.25)
            t[flips] = (t[flips] + rng.integers(1, 4, int(flips.sum()))) % 4
            if mode == 2 and tlen > 4:
                cut = int(rng.integers(1, tlen - 1))
                t = np.concatenate([t[:cut], t[cut + 1:], rng.integers(0, 4, 1)]).astype(np.uint8)
        if trial % 7 == 0:
            q[int(rng.integers(0, qlen))] = 4
        h0 = int(rng.integers(1, 40))
        w = int(rng.integers(1, 30))
        kw = dict(o_del=int(rng.integers(0, 8)), e_del=int(rng.integers(1, 3)), o_ins=int(rng.integers(0, 8)),
                  e_ins=int(rng.integers(1, 3)), zdrop=int(rng.choice([0, 5, 100])), end_bonus=5)
        r, cells = ext(O, q, t, h0, w=w, variant=variant, **kw)
        m = MM.extend(mat, q, t, h0, w, variant=variant, **kw)
        mm_cells = m.pop("cells")
        assert r == m, (trial, variant, r, m)
        assert cells == mm_cells


def test_chain2aln_perfect_read(O, B):
    """Clean read: left flank 30, seed 40, right flank 31 -> end-to-end alignment, score = read length."""
    from helpers import oracle_chain2aln
    rng = np.random.default_rng(11)
    read = rng.integers(0, 4, 101).astype(np.uint8)
    ql = read[:30][::-1].copy(); qr = read[70:].copy()
    tl = np.concatenate([ql, rng.integers(0, 4, 25)]).astype(np.uint8)
    tr = np.concatenate([qr, rng.integers(0, 4, 26)]).astype(np.uint8)
    seeds = [dict(q_left=ql, q_right=qr, t_left=tl, t_right=tr, init_score=-1, qbeg=30, h0=40, id=77)]
    out, _ = oracle_chain2aln(O, B, B.make_params2(), seeds)
    r = out[0]
    assert (r["id"], r["qb"], r["qe"], r["rb"], r["re"], r["score"], r["truesc"], r["w"]) == (77, 0, 31, -30, 31, 101, 101, 100)
