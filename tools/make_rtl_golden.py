#!/usr/bin/env python3
"""Freeze outputs of the REFERENCE ITSELF into tests/golden/rtl_*.npz.

Runs only in the container where /root/reference is mounted: `make -C oracle` translates the Verilog there into
oracle/_ref/librtlsim.so (oracle/rtl2c/v2c.py, mechanical; oracle/rtl2c/tb.cpp, test bench), this script drives that
cycle model with seeded tasks and stores inputs + the RTL's outputs.  Nothing here asks the C oracle or the CUDA path
for an answer -- the files hold what the mounted RTL computes; tests/test_rtl_pin.py then checks the oracle (CPU suite)
and the CUDA path (GPU suite) against them.

    python tools/make_rtl_golden.py            # ~2-3 minutes

Files (all inputs are regenerated from the seeds below; every array is stored so the tests never need the RTL):
  rtl_sw_extend.npz        level 1: sw_pe_array_sw_extend alone, in-envelope tasks (scores <= 127, qlen <= 255, w <= 63)
  rtl_sw_extend_wide.npz   level 1: tasks that leave the 8-bit envelope (for the RTL-width model only)
  rtl_pe_array.npz         level 3: whole sw_pe_array batches, TBB image in -> RBB image out
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rtlsim as R  # noqa: E402
from oracle import rtl_envelope  # noqa: E402  (a predicate on the inputs only; no oracle arithmetic is used here)

GOLD = os.path.join(ROOT, "tests", "golden")


def max_gap(qlen, end_bonus, o, e, a=1):
    """ksw_extend2's max_ins / max_del, which the FPGA receives precomputed (sw_pe_array_proc_element.v:924-934)."""
    return max(1, int((qlen * a + end_bonus - o) / e + 1.0))


def mutate(rng, q, sub, indel, tail):
    out = []
    for b in q:
        r = rng.random()
        if r < indel / 2:
            continue
        if r < indel:
            out.append(int(rng.integers(0, 4)))
        out.append(int((b + rng.integers(1, 4)) % 4) if (rng.random() < sub and b < 4) else int(b))
    out += [int(x) for x in rng.integers(0, 4, tail)]
    return np.array(out, dtype=np.uint8)


def level1_tasks(rng, n, wide):
    tasks = []
    gaps = [(6, 1, 6, 1), (6, 1, 6, 1), (4, 2, 5, 1), (5, 3, 7, 2), (0, 1, 0, 1), (10, 1, 3, 2)]
    k = -1
    while len(tasks) < n:
        k += 1
        o_ins, e_ins, o_del, e_del = gaps[k % len(gaps)]
        if wide:
            qlen = int(rng.integers(1, 256))
            h0 = int(rng.integers(1, 256))
            w = int(rng.integers(1, 128))
        else:
            qlen = int(rng.integers(1, 121))
            h0 = int(rng.integers(1, 128 - qlen))               # h0 + qlen*max(mat) <= 127: the 8-bit datapath is exact
            w = int(rng.choice([1, 2, 3, 5, 8, 15, 30, 50, 63]))
        q = rng.integers(0, 4, qlen).astype(np.uint8)
        mode = k % 7
        if mode == 0:
            t = rng.integers(0, 4, int(rng.integers(1, qlen + 40))).astype(np.uint8)          # unrelated
        elif mode == 1:
            q = rng.integers(0, 2, qlen).astype(np.uint8)                                        # low complexity: ties
            t = mutate(rng, q, 0.05, 0.02, int(rng.integers(0, 30)))
        else:
            sub, indel = [(0.0, 0.0), (0.01, 0.0), (0.04, 0.01), (0.1, 0.05), (0.02, 0.08)][mode - 2]
            t = mutate(rng, q, sub, indel, int(rng.integers(0, 60)))
            if rng.random() < 0.3:
                t = t[: int(rng.integers(1, len(t) + 1))]                                       # tlen < qlen happens
        if len(t) == 0:
            t = np.zeros(1, np.uint8)
        if rng.random() < 0.15:
            q[int(rng.integers(0, qlen))] = 4                                                    # N in the query
        if rng.random() < 0.1:
            t[int(rng.integers(0, len(t)))] = 4
        end_bonus = int(rng.choice([0, 5, 5, 9]))
        if rtl_envelope(qlen, len(t), h0, w, o_del, e_del, o_ins, e_ins) == wide:
            continue                                             # narrow file: inside the envelope; wide file: outside
        tasks.append(dict(q=q, t=t, h0=h0, w=w, o_ins=o_ins, e_ins=e_ins, o_del=o_del, e_del=e_del,
                          reg_score=int(rng.integers(0, 128)), end_bonus=end_bonus, max_ins=max_gap(qlen, end_bonus, o_ins, e_ins),
                          max_del=max_gap(qlen, end_bonus, o_del, e_del)))
    return tasks


def run_level1(tasks, name):
    n = len(tasks)
    out = np.zeros((n, 7), np.int32)
    clocks = np.zeros(n, np.int64)
    for i, t in enumerate(tasks):
        out[i], clocks[i] = R.sw_extend(t["q"], t["t"], t["h0"], t["w"], t["o_ins"], t["e_ins"], t["o_del"], t["e_del"],
                                        t["reg_score"], t["max_ins"], t["max_del"], scramble_seed=i + 1)
    qoff = np.zeros(n + 1, np.int64)
    toff = np.zeros(n + 1, np.int64)
    np.cumsum([len(t["q"]) for t in tasks], out=qoff[1:])
    np.cumsum([len(t["t"]) for t in tasks], out=toff[1:])
    scal = np.array([[t[k] for k in ("h0", "w", "o_ins", "e_ins", "o_del", "e_del", "reg_score", "max_ins", "max_del", "end_bonus")]
                     for t in tasks], np.int32)
    np.savez_compressed(os.path.join(GOLD, name), qbuf=np.concatenate([t["q"] for t in tasks]), qoff=qoff,
                        tbuf=np.concatenate([t["t"] for t in tasks]), toff=toff, scalars=scal, rtl=out, clocks=clocks,
                        scalar_names=np.array("h0 w o_ins e_ins o_del e_del reg_score max_ins max_del end_bonus".split()),
                        rtl_names=np.array(R.EXT_FIELDS))
    print(f"{name}: {n} tasks, {int(clocks.sum())} clocks, second band try on {(out[:, 1] != scal[:, 1]).sum()}")


# ---------------------------------------------------------------------------------------------------------------------
def build_tbb(hdr, seeds):
    """TBB image written from the RTL's field map alone: header words 0-2
    (sw_pe_array_proc_element.v:815-820,915-918; sw_pe_array_task_parse.v:944), 8 parameter words per task
    (proc_element.v:880-892,871-874,826-828,924-934,807), data offset in words (task_parse.v:1924-1936), bases 4 bit
    each, first base in bits 31:28 (proc_element.v:1638,1677), segments qL(rev) qR tL(rev) tR back to back."""
    n = len(seeds)
    tbb = np.zeros(65536, np.uint32)
    tbb[0] = hdr["o_del"] | (hdr["e_del"] << 8) | (hdr["o_ins"] << 16) | (hdr["e_ins"] << 24)
    tbb[1] = hdr["pen_clip5"] | (hdr["pen_clip3"] << 8) | (hdr["w"] << 16)
    tbb[2] = n
    data = 8 + 8 * n
    off = 0
    for i, s in enumerate(seeds):
        p = 8 + 8 * i
        ql, qr, tl, tr = (len(s[k]) for k in ("q_left", "q_right", "t_left", "t_right"))
        tbb[p + 0] = ql | (tl << 16)
        tbb[p + 1] = qr | (tr << 16)
        tbb[p + 2] = off + 77                            # only differences to task 0's offset matter (task_parse.v:1928)
        tbb[p + 3] = (s["init_score"] & 0xFFFF) | (s["qbeg"] << 16)
        tbb[p + 4] = s["h0"] & 0xFF
        tbb[p + 5] = max_gap(ql, hdr["pen_clip5"], hdr["o_ins"], hdr["e_ins"]) | (max_gap(ql, hdr["pen_clip5"], hdr["o_del"], hdr["e_del"]) << 16)
        tbb[p + 6] = max_gap(qr, hdr["pen_clip3"], hdr["o_ins"], hdr["e_ins"]) | (max_gap(qr, hdr["pen_clip3"], hdr["o_del"], hdr["e_del"]) << 16)
        tbb[p + 7] = s["id"]
        bases = np.concatenate([s["q_left"], s["q_right"], s["t_left"], s["t_right"]]).astype(np.uint32)
        nw = (len(bases) + 7) // 8
        padded = np.zeros(nw * 8, np.uint32)
        padded[: len(bases)] = bases
        words = np.zeros(nw, np.uint32)
        for k in range(8):
            words |= padded[k::8] << np.uint32(28 - 4 * k)
        assert data + off + nw <= 65536, "batch does not fit the TBB"
        tbb[data + off: data + off + nw] = words
        off += nw
    return tbb


def seed_tasks(rng, n, read_len, sub, indel, id0, tiny_h0=False):
    """Seeds of synthetic reads: left flank reversed, right flank forward, reference windows with slack (mem_chain2aln)."""
    seeds = []
    for k in range(n):
        L = int(rng.integers(30, read_len + 1))
        slen = int(rng.integers(15, min(40, L) + 1))
        sbeg = int(rng.integers(0, L - slen + 1))
        read = rng.integers(0, 4, L).astype(np.uint8)
        if rng.random() < 0.1:
            read[int(rng.integers(0, L))] = 4
        lq = read[:sbeg][::-1].copy()
        rq = read[sbeg + slen:].copy()
        unrelated = rng.random() < 0.08
        lt = mutate(rng, lq if not unrelated else rng.integers(0, 4, len(lq)).astype(np.uint8), sub, indel, int(rng.integers(0, 12))) if len(lq) else np.zeros(0, np.uint8)
        rt = mutate(rng, rq, sub, indel, int(rng.integers(0, 12))) if len(rq) else np.zeros(0, np.uint8)
        mode = k % 9
        if mode == 7:
            lq, lt = lq[:0], lt[:0]
        if mode == 8:
            rq, rt = rq[:0], rt[:0]
        if len(lq) and not len(lt):
            lt = np.zeros(1, np.uint8)
        if len(rq) and not len(rt):
            rt = np.zeros(1, np.uint8)
        h0 = int(rng.integers(1, 4)) if tiny_h0 else slen         # tiny h0: where the FPGA's carried second try shows
        seeds.append(dict(q_left=lq, q_right=rq, t_left=lt, t_right=rt, init_score=h0, qbeg=len(lq), h0=h0,
                          id=id0 + k))
    return seeds


def run_level3():
    rng = np.random.default_rng(2026)
    batches = []
    plans = [
        # (tasks, read length, sub, indel, header)
        (1, 60, 0.0, 0.0, dict(o_del=6, e_del=1, o_ins=6, e_ins=1, pen_clip5=5, pen_clip3=5, w=50)),
        (0, 60, 0.0, 0.0, dict(o_del=6, e_del=1, o_ins=6, e_ins=1, pen_clip5=5, pen_clip3=5, w=50)),
        (19, 101, 0.01, 0.0, dict(o_del=6, e_del=1, o_ins=6, e_ins=1, pen_clip5=5, pen_clip3=5, w=50)),
        (21, 101, 0.03, 0.01, dict(o_del=6, e_del=1, o_ins=6, e_ins=1, pen_clip5=5, pen_clip3=5, w=100)),
        (200, 101, 0.02, 0.005, dict(o_del=6, e_del=1, o_ins=6, e_ins=1, pen_clip5=5, pen_clip3=5, w=100)),
        (300, 90, 0.05, 0.03, dict(o_del=6, e_del=1, o_ins=6, e_ins=1, pen_clip5=5, pen_clip3=9, w=20)),
        (300, 80, 0.04, 0.04, dict(o_del=4, e_del=2, o_ins=5, e_ins=1, pen_clip5=0, pen_clip3=7, w=8)),
        (819, 70, 0.02, 0.01, dict(o_del=6, e_del=1, o_ins=6, e_ins=1, pen_clip5=5, pen_clip3=5, w=40)),
        (400, 101, 0.01, 0.06, dict(o_del=6, e_del=1, o_ins=6, e_ins=1, pen_clip5=5, pen_clip3=5, w=2)),     # band retries
        (400, 101, 0.02, 0.08, dict(o_del=3, e_del=1, o_ins=3, e_ins=1, pen_clip5=5, pen_clip3=5, w=3)),
        # tiny h0 + noisy flanks: the second band try scores BELOW the first, so the FPGA's carried maxima (SURVEY
        # appendix C row 5) and ksw_extend2's fresh start give different records
        (800, 101, 0.3, 0.1, dict(o_del=6, e_del=1, o_ins=6, e_ins=1, pen_clip5=5, pen_clip3=5, w=17), True),
        (800, 101, 0.4, 0.2, dict(o_del=6, e_del=1, o_ins=6, e_ins=1, pen_clip5=5, pen_clip3=5, w=8), True),
    ]
    tbbs, rbbs, nwords, clocks = [], [], [], []
    for b, plan in enumerate(plans):
        n, L, sub, indel, hdr = plan[:5]
        seeds = seed_tasks(rng, n, L, sub, indel, id0=(b + 1) * 100000, tiny_h0=len(plan) > 5)
        tbb = build_tbb(hdr, seeds)
        rbb, nw, clk = R.pe_array(tbb, scramble_seed=b + 1)
        assert nw == 5 * n, (nw, n)
        print(f"rtl_pe_array batch {b}: {n} tasks, {clk} clocks")
        tbbs.append(tbb); rbbs.append(rbb); nwords.append(nw); clocks.append(clk)
        batches.append(seeds)
    np.savez_compressed(os.path.join(GOLD, "rtl_pe_array.npz"), tbb=np.array(tbbs), rbb=np.array(rbbs),
                        n_words=np.array(nwords), clocks=np.array(clocks))


# Known-answer tasks on which the two recurrence policies differ (tests/test_oracle_kat.py derives the answers by hand):
# the RTL must give the V1 answer.  reg_score = the expected V1 score, so that sw_extend stops after one band try.
KATS = [
    dict(name="no_zero_guard", q=[0, 0], t=[0], h0=1, w=1, v1=(2, 1, 1, 1, 1, 0), v2=(2, 1, 1, 1, 0, 0)),
    dict(name="first_column_unconditional", q=[0, 0, 0], t=[1, 1, 1, 0], h0=15, w=1, v1=(15, 0, 0, 4, 5, 0), v2=(15, 0, 0, 3, 3, 0)),
    dict(name="narrowing_run_around_mj", q=[0, 0], t=[1, 1, 0, 0], h0=9, w=2, v1=(9, 0, 0, 2, 1, 0), v2=(9, 0, 0, 4, 3, 0)),
]


def run_kats():
    rows = []
    for k in KATS:
        q, t = np.array(k["q"], np.uint8), np.array(k["t"], np.uint8)
        mi = max_gap(len(q), 5, 6, 1)
        out, _ = R.sw_extend(q, t, k["h0"], k["w"], reg_score=k["v1"][0], max_ins=mi, max_del=mi, scramble_seed=7)
        rows.append(out)
        print("kat", k["name"], "rtl", out)
    np.savez_compressed(os.path.join(GOLD, "rtl_kat.npz"), names=np.array([k["name"] for k in KATS]), rtl=np.array(rows, np.int32))


if __name__ == "__main__":
    if not R.available():
        sys.exit("oracle/_ref/librtlsim.so missing: run `make -C oracle` where /root/reference is mounted")
    os.makedirs(GOLD, exist_ok=True)
    run_kats()
    if "kat" in sys.argv[1:]:
        sys.exit(0)
    if "l3" not in sys.argv[1:]:
        run_level1(level1_tasks(np.random.default_rng(11), 6000, wide=False), "rtl_sw_extend.npz")
        run_level1(level1_tasks(np.random.default_rng(12), 3000, wide=True), "rtl_sw_extend_wide.npz")
    run_level3()
