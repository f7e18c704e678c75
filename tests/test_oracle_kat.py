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


# ---------------------------------------------------------------------------------------------------------------------
# Known answers on which V1 (what the mounted RTL computes) and V2 (upstream BWA) DIFFER, derived by hand below, one
# per policy point; the translated RTL's own outputs for the same tasks are frozen in tests/golden/rtl_kat.npz
# (tools/make_rtl_golden.py) and must equal the V1 column.  Defaults: a=1 b=4 o=6 e=1, zdrop off, end_bonus 5.
# The fourth point (gap open taken from h instead of M, sw_pe_array_sw_extend.v:1866,1863) cannot change an output under
# this scoring -- a gap opened from an E- or F-derived h means an insertion right next to a deletion, 2(o+e) = 14
# against 5 for the mismatch it replaces -- and an exhaustive search over all binary tasks up to 5 x 6 bases finds no
# difference from it alone; it is covered by the randomised V1/V2 comparisons.
V1V2_KATS = [
    # (1) no "M ? M+s : 0" guard (sx:1797).  q = AA, t = A, h0 = 1, w = 1.  First row eh.h = [1, 0, 0].  Row 0, h1 = max(0,1-7) = 0:
    #     j=0: M = 1+1 = 2 = h (m = 2, mj = 0);  j=1: the stored H is 0: V1 h = 0+1 = 1, V2 h = 0.  The row reaches qlen, so
    #     gscore = h of the last cell: 1 (V1) / 0 (V2), gtle = 1.  score = 2 at (qle, tle) = (1, 1).
    dict(name="no_zero_guard", q=[0, 0], t=[0], h0=1, w=1, v1=(2, 1, 1, 1, 1, 0), v2=(2, 1, 1, 1, 0, 0)),
    # (2) first-column value applied unconditionally (sx:1795-1796).  q = AAA, t = CCCA, h0 = 15, w = 1.  Rows 0-2 only see
    #     mismatches: row maxima 11, 7, 3 (never above h0 = 15, so score = 15, qle = tle = 0); row 1 reaches qlen with h = 0
    #     (gscore 0), row 2 with h = 3 (gscore 3, max_ie = 2) and narrows to beg = 3 = qlen.  Row 3 is empty; its h1 is
    #     max(0, 15-(6+4)) = 5 in V1 whatever beg is, 0 in V2 because beg != 0; the empty row still performs the
    #     "j == qlen" update with that h1: V1 gscore = 5 at gtle = 4, V2 keeps gscore = 3 at gtle = 3.
    dict(name="first_column_unconditional", q=[0, 0, 0], t=[1, 1, 1, 0], h0=15, w=1, v1=(15, 0, 0, 4, 5, 0), v2=(15, 0, 0, 3, 3, 0)),
    # (3) narrowing = non-zero run around mj (sx:1766-1769,1779) vs BWA's zero scan.  q = AA, t = CCAA, h0 = 9, w = 2.
    #     Row 0: H = [5, 0]; row 1 (h1 = 1): H = [0, 1], mj = 1, gscore = 1 at gtle = 2.  V1 scans left from mj: the stored
    #     H of column 1 is 0, so beg = 2 = qlen and row 2 is empty: m = 0, stop -> (9,0,0,2,1,0).  V2 only skips leading
    #     cells with h = e = 0 from beg: the first-column value 1 is non-zero, beg stays 0; rows 2 and 3 then match
    #     (H = [2, 0] and [., 3]): gscore = 3 at gtle = 4 -> (9,0,0,4,3,0).
    dict(name="narrowing_run_around_mj", q=[0, 0], t=[1, 1, 0, 0], h0=9, w=2, v1=(9, 0, 0, 2, 1, 0), v2=(9, 0, 0, 4, 3, 0)),
]
FIELDS = ("score", "qle", "tle", "gtle", "gscore", "max_off")


@pytest.mark.parametrize("kat", V1V2_KATS, ids=[k["name"] for k in V1V2_KATS])
def test_v1_and_v2_differ_as_derived(O, kat):
    p = O.make_params(zdrop=0)
    q, t = np.array(kat["q"], np.uint8), np.array(kat["t"], np.uint8)
    for variant, want in ((1, kat["v1"]), (2, kat["v2"])):
        r, _ = O.extend_one(p, q, t, kat["h0"], kat["w"], variant=variant)
        assert tuple(int(r[f]) for f in FIELDS) == want, f"V{variant}"
    assert kat["v1"] != kat["v2"]


def test_the_rtl_gives_the_v1_answers(O):
    """Frozen outputs of the translated RTL (score, aw, qle, tle, gtle, gscore, max_off), and a live run where it is built."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rtl_kat.npz"))
    assert [str(n) for n in g["names"]] == [k["name"] for k in V1V2_KATS]
    from oracle import rtlsim as R
    for kat, row in zip(V1V2_KATS, g["rtl"]):
        assert (int(row[0]),) + tuple(int(x) for x in row[2:]) == kat["v1"] and int(row[1]) == kat["w"]
        if R.available():
            q, t = np.array(kat["q"], np.uint8), np.array(kat["t"], np.uint8)
            gap = max(1, int((len(q) + 5 - 6) / 1 + 1.0))
            live, _ = R.sw_extend(q, t, kat["h0"], kat["w"], reg_score=kat["v1"][0], max_ins=gap, max_del=gap)
            assert np.array_equal(live, row)


def test_matrix_model_agrees_on_the_v1_v2_kats():
    from oracle import matrix_model as MM
    import oracle
    mat = oracle.bwa_fill_scmat(1, 4)
    for kat in V1V2_KATS:
        for variant, want in ((1, kat["v1"]), (2, kat["v2"])):
            r = MM.extend(mat, kat["q"], kat["t"], kat["h0"], kat["w"], zdrop=0, variant=variant)
            assert tuple(r[f] for f in FIELDS) == want
