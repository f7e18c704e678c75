"""FPGA wire formats: the TBB image bsw_tbb_encode builds, parsed back by an independent
Python reader written from the RTL field map, and the RBB record decoder."""
import numpy as np
import pytest


from helpers import parse_tbb  # noqa: E402


def test_nibble_order_example(B):
    """Appendix A.1: bases A C G T A C G N pack to 0x01230124."""
    q = np.array([0, 1, 2, 3, 0, 1, 2, 4], np.uint8)
    e = np.zeros(0, np.uint8)
    tbb = B.tbb_encode(B.make_params2(), [dict(q_left=e, q_right=q, t_left=e, t_right=e, init_score=5, qbeg=0, h0=5, id=9)])
    assert int(tbb[2]) == 1 and int(tbb[8 + 8]) == 0x01230124


def test_tbb_roundtrip(B):
    from helpers import seeds_from_flat
    t = B.synth_tasks("cfg1_101bp", 400, seed=4)
    seeds = seeds_from_flat(t, 200, unset_score_every=5)
    P2 = B.make_params2(w=100, pen_clip5=5, pen_clip3=7)
    tbb = B.tbb_encode(P2, seeds)
    hdr, tasks = parse_tbb(tbb)
    assert hdr == dict(o_del=6, e_del=1, o_ins=6, e_ins=1, pen_clip5=5, pen_clip3=7, w=100, n=200)
    for s, d in zip(seeds, tasks):
        assert d["qlen"] == [len(s["q_left"]), len(s["q_right"])] and d["tlen"] == [len(s["t_left"]), len(s["t_right"])]
        assert d["init_score"] == (s["init_score"] & 0xffff) and d["qbeg"] == s["qbeg"] and d["h0"] == s["h0"] and d["id"] == s["id"]
        want = list(s["q_left"]) + list(s["q_right"]) + list(s["t_left"]) + list(s["t_right"])
        assert d["bases"] == [int(x) for x in want]
        for side, eb in ((0, 5), (1, 7)):       # ksw_extend2: max_gap = (qlen*a + end_bonus - o)/e + 1, at least 1
            exp = max(1, int((d["qlen"][side] * 1 + eb - 6) / 1 + 1.0))
            assert d["max_ins"][side] == exp and d["max_del"][side] == exp


def test_tbb_limits(B):
    e = np.zeros(0, np.uint8)
    big = np.zeros(256, np.uint8)
    with pytest.raises(B.BswError) as err:      # qlen is an 8-bit field (proc_element.v:880)
        B.tbb_encode(B.make_params2(), [dict(q_left=e, q_right=big, t_left=e, t_right=big, init_score=5, qbeg=0, h0=5)])
    assert err.value.code == B.BSW_EWIRE
    small = np.zeros(10, np.uint8)
    many = [dict(q_left=e, q_right=small, t_left=e, t_right=small, init_score=5, qbeg=0, h0=5)] * 820
    with pytest.raises(B.BswError):             # 4096/5 = 819 records fit an RBB (fill_resulBuf.v:378)
        B.tbb_encode(B.make_params2(), many)


def test_rbb_decode(B):
    rbb = np.zeros(B.RBB_WORDS, np.uint32)
    # [id][qe<<16|qb][re<<16|rb][truesc<<16|score][w]   (proc_element.v:1187-1205,1662-1665)
    rbb[0:5] = [42, (31 << 16) | 0, (31 << 16) | ((-30) & 0xffff), (101 << 16) | 101, 100]
    rbb[5:10] = [43, (20 << 16) | 3, (25 << 16) | ((-7) & 0xffff), (55 << 16) | 60, 200]
    r = B.rbb_decode(rbb, 2)
    assert tuple(r[0]) == (42, 0, 31, -30, 31, 101, 101, 100)
    assert tuple(r[1]) == (43, 3, 20, -7, 25, 60, 55, 200)
