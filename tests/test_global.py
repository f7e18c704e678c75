"""SURVEY 8 f.4 -- banded global alignment with traceback (ksw_global2), the DP that follows seed extension in BWA-MEM.
The algorithm is not in /root/reference (the FPGA stops at the extension); oracle.global_align restates the published BWA
routine, so this row's parity is UNPINNED.  What the CPU tests hold instead are properties: the CIGAR consumes exactly the
query and the target, re-scoring the CIGAR gives the returned score, and with a full band the score equals an
independent Gotoh (numpy, three matrices) written from the textbook recurrence."""
import numpy as np
import pytest


MINF = -(1 << 29)


def gotoh(q, t, mat, o_ins, e_ins, o_del, e_del, w=None):
    """Three-matrix global alignment, written from the textbook recurrence; w: only cells with |i - j| <= w exist."""
    n, m = len(q), len(t)
    if w is None: w = n + m
    H = np.full((m + 1, n + 1), MINF, dtype=np.int64); E = H.copy(); F = H.copy()
    H[0, 0] = 0
    for j in range(1, min(n, w) + 1): H[0, j] = -(o_ins + e_ins * j)
    for i in range(1, min(m, w) + 1): H[i, 0] = -(o_del + e_del * i)
    for i in range(1, m + 1):
        for j in range(max(1, i - w), min(n, i + w) + 1):
            E[i, j] = max(E[i - 1, j] - e_del, H[i - 1, j] - o_del - e_del)      # gap in the query (deletion)
            F[i, j] = max(F[i, j - 1] - e_ins, H[i, j - 1] - o_ins - e_ins)      # gap in the target (insertion)
            H[i, j] = max(H[i - 1, j - 1] + mat[t[i - 1] * 5 + q[j - 1]], E[i, j], F[i, j])
    return int(H[m, n])


def rescore(cigar, q, t, mat, o_ins, e_ins, o_del, e_del):
    i = j = 0; s = 0
    for c in cigar:
        op, ln = int(c) & 15, int(c) >> 4
        if op == 0:
            for k in range(ln): s += int(mat[t[i + k] * 5 + q[j + k]])
            i += ln; j += ln
        elif op == 1: s -= o_ins + e_ins * ln; j += ln
        else: s -= o_del + e_del * ln; i += ln
    return s, j, i


def cases(seed, n, lo=1, hi=60):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        ql = int(rng.integers(lo, hi)); q = rng.integers(0, 4, ql).astype(np.uint8)
        t = list(q)
        for _ in range(int(rng.integers(0, 4))):                  # a few edits
            kind = rng.integers(0, 3); pos = int(rng.integers(0, max(1, len(t))))
            if kind == 0 and t: t[pos] = (t[pos] + 1) % 4
            elif kind == 1: t[pos:pos] = list(rng.integers(0, 4, int(rng.integers(1, 6))))
            elif len(t) > 6: del t[pos:pos + int(rng.integers(1, 5))]
        if not t: t = [0]
        if rng.random() < 0.1: q[int(rng.integers(0, ql))] = 4     # an N
        out.append((q, np.array(t, dtype=np.uint8)))
    return out


def test_known_answer(O):
    p = O.make_params()
    q = np.array([0, 1, 2, 3, 0, 1, 2, 3, 0, 1], dtype=np.uint8)
    t = np.array([0, 1, 2, 2, 3, 0, 1, 2, 3, 0, 1], dtype=np.uint8)         # one inserted target base -> 1D
    score, cig = O.global_align(p, q, t, 5)
    assert score == 10 - 7                                                    # 10 matches, one deletion of length 1 (6 + 1)
    assert [(int(c) >> 4, int(c) & 15) for c in cig] in ([(3, 0), (1, 2), (7, 0)], [(2, 0), (1, 2), (8, 0)])


def test_cigar_consumes_both_and_rescoring_gives_the_score(O):
    p = O.make_params()
    mat = np.frombuffer(bytes(p.mat), dtype=np.int8)
    for q, t in cases(11, 300):
        w = abs(len(q) - len(t)) + 8
        score, cig = O.global_align(p, q, t, w)
        s, qc, tc = rescore(cig, q, t, mat, p.o_ins, p.e_ins, p.o_del, p.e_del)
        assert (qc, tc) == (len(q), len(t))
        assert s == score


def test_full_band_equals_gotoh(O):
    p = O.make_params()
    mat = np.frombuffer(bytes(p.mat), dtype=np.int8)
    for q, t in cases(12, 120, 1, 40):
        score, _ = O.global_align(p, q, t, len(q) + len(t))
        assert score == gotoh(q, t, mat, p.o_ins, p.e_ins, p.o_del, p.e_del)


def test_banded_score_equals_banded_gotoh(O):
    """The band of ksw_global2 is |i - j| <= w on the (target row, query column) grid: an independent banded Gotoh gives
    the same optimum for every band that can reach the last cell."""
    p = O.make_params(o_del=5, e_del=2, o_ins=7, e_ins=1)
    mat = np.frombuffer(bytes(p.mat), dtype=np.int8)
    rng = np.random.default_rng(13)
    for q, t in cases(13, 150, 1, 50):
        w = abs(len(q) - len(t)) + int(rng.integers(0, 6))
        score, cig = O.global_align(p, q, t, w)
        assert score == gotoh(q, t, mat, p.o_ins, p.e_ins, p.o_del, p.e_del, w)
        s, qc, tc = rescore(cig, q, t, mat, p.o_ins, p.e_ins, p.o_del, p.e_del)
        assert (s, qc, tc) == (score, len(q), len(t))


def test_max_cigar_overflow_is_reported(O):
    p = O.make_params()
    q = np.array([0, 1] * 20, dtype=np.uint8); t = np.array([0, 1] * 10 + [2] * 6 + [0, 1] * 10, dtype=np.uint8)
    score, cig = O.global_align(p, q, t, 10, max_cigar=2)
    assert cig is None
    score2, cig2 = O.global_align(p, q, t, 10)
    assert score2 == score and len(cig2) == 3


@pytest.mark.gpu
def test_cuda_global_equals_oracle(B, O, ctx):
    p = O.make_params(); pb = B.make_params()
    cs = cases(21, 2000, 1, 200) + cases(22, 40, 300, 900)
    ws = [abs(len(q) - len(t)) + int(k % 3) * 20 + 1 for k, (q, t) in enumerate(cs)]
    score, cigs = ctx.global_batch(pb, [c[0] for c in cs], [c[1] for c in cs], ws)
    for k, (q, t) in enumerate(cs):
        s, cig = O.global_align(p, q, t, ws[k])
        assert s == score[k], k
        assert np.array_equal(cig, cigs[k]), k


@pytest.mark.gpu
def test_cuda_global_errors(B, ctx):
    pb = B.make_params()
    q = np.array([0, 1] * 20, dtype=np.uint8); t = np.array([0, 1] * 10 + [2] * 6 + [0, 1] * 10, dtype=np.uint8)
    with pytest.raises(B.BswError) as e:
        ctx.global_batch(pb, [q], [t], [10], max_ops=2)
    assert e.value.code == B.BSW_ERANGE
    with pytest.raises(B.BswError) as e:
        ctx.global_batch(pb, [q], [t], [2])                                  # band narrower than the length difference
    assert e.value.code == B.BSW_EINVAL
