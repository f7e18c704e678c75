"""Host task builder: bsw_chain_window / bsw_build_seed_tasks / bsw_finish_seed, driven by a C program
compiled against include/bsw.h (tests/c/chain_caller.c), checked against a Python restatement of mem_chain2aln's
bookkeeping; on the GPU the same program runs the level-2 batch and its records are compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def caller(built, tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("c") / "chain_caller")
    lib_dir = os.path.join(ROOT, "bwa-mem-sw_b200")
    subprocess.check_call(["gcc", "-O1", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "chain_caller.c"), "-o", exe, "-L", lib_dir, "-lbsw", f"-Wl,-rpath,{lib_dir}"])
    return exe


def cal_max_gap(qlen, w, a=1, o=6, e=1):
    l = max(int((qlen * a - o) / e + 1.0), 1)
    return min(l, w << 1)


def make_case(seed, nseeds, l_query=101, l_pac=3000, w=100):
    rng = np.random.default_rng(seed)
    ref = rng.integers(0, 4, l_pac).astype(np.uint8)
    pos = int(rng.integers(300, l_pac - 400))
    read = ref[pos:pos + l_query].copy()
    flips = rng.random(l_query) < 0.04
    read[flips] = (read[flips] + rng.integers(1, 4, int(flips.sum()))) % 4
    seeds = []
    for k in range(nseeds):
        ln = int(rng.integers(15, 40))
        qb = 0 if k == 0 else (l_query - ln if k == 1 else int(rng.integers(0, l_query - ln + 1)))
        seeds.append((pos + qb, qb, ln))
    if nseeds > 2:
        seeds[2] = (pos, 0, l_query)                       # a seed that spans the whole read: no flank at all
    return read, ref, seeds, l_query, l_pac, w


def run(caller, mode, case):
    read, ref, seeds, l_query, l_pac, w = case
    text = f"{l_query} {l_pac} {len(seeds)} {w}\n" + "".join(map(str, read)) + "\n" + "".join(map(str, ref)) + "\n" + \
           "".join(f"{rb} {qb} {ln}\n" for rb, qb, ln in seeds)
    out = subprocess.run([caller, mode], input=text, capture_output=True, text=True, check=True).stdout.split("\n")
    return [l.split() for l in out if l]


def expected_tasks(case):
    read, ref, seeds, l_query, l_pac, w = case
    b = min(rb - (qb + cal_max_gap(qb, w)) for rb, qb, ln in seeds)
    e = max(rb + ln + ((l_query - qb - ln) + cal_max_gap(l_query - qb - ln, w)) for rb, qb, ln in seeds)
    r0, r1 = max(b, 0), min(e, 2 * l_pac)
    if r0 < l_pac < r1:
        r1 = l_pac
    tasks = []
    for rb, qb, ln in seeds:
        ql, tl = read[:qb][::-1], ref[r0:rb][::-1]
        qr, tr = read[qb + ln:], ref[rb + ln:r1]
        tasks.append(dict(q_left=ql, t_left=tl if qb else ql[:0], q_right=qr, t_right=tr if len(qr) else qr[:0],
                          init_score=-1 if qb else ln, qbeg=qb, h0=ln))
    return (r0, r1), tasks


def digits(a):
    return "".join(map(str, a)) if len(a) else "-"


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_seed_tasks_from_a_chain(caller, seed):
    case = make_case(seed, 6)
    lines = run(caller, "host", case)
    (r0, r1), want = expected_tasks(case)
    assert lines[0] == ["rmax", str(r0), str(r1)]
    body = lines[1:]
    for i, t in enumerate(want):
        hdr, ql, tl, qr, tr = body[5 * i: 5 * i + 5]
        assert hdr == ["task", str(i), "qlen", str(len(t["q_left"])), str(len(t["q_right"])), "tlen", str(len(t["t_left"])),
                       str(len(t["t_right"])), "init", str(t["init_score"]), "qbeg", str(t["qbeg"]), "h0", str(t["h0"])]
        assert ql[1] == digits(t["q_left"]) and tl[1] == digits(t["t_left"]) and qr[1] == digits(t["q_right"]) and tr[1] == digits(t["t_right"])


@pytest.mark.gpu
def test_c_caller_end_to_end(caller, B, O):
    from helpers import oracle_chain2aln
    case = make_case(7, 40)
    lines = run(caller, "gpu", case)
    (r0, r1), want = expected_tasks(case)
    seeds = [dict(t, id=i) for i, t in enumerate(want)]
    P2 = B.make_params2(B.make_params(zdrop=100, end_bonus=5), w=case[5], pen_clip5=5, pen_clip3=5)
    rec, _ = oracle_chain2aln(O, B, P2, seeds)
    recs = [l for l in lines if l[0] == "rec"]
    alns = [l for l in lines if l[0] == "aln"]
    assert len(recs) == len(seeds) == len(alns)
    l_query = case[3]
    for i, (r, a) in enumerate(zip(recs, alns)):
        assert [int(x) for x in r[1:]] == [int(rec[i][f]) for f in ("id", "qb", "qe", "rb", "re", "score", "truesc", "w")]
        rb, qb, ln = case[2][i]
        exp_qb = int(rec[i]["qb"]) if qb else 0
        exp_rb = rb + int(rec[i]["rb"]) if qb else rb
        exp_qe = qb + ln + int(rec[i]["qe"]) if qb + ln != l_query else l_query
        exp_re = rb + ln + int(rec[i]["re"]) if qb + ln != l_query else rb + ln
        sc, tsc = int(rec[i]["score"]), int(rec[i]["truesc"])
        if qb == 0 and qb + ln == l_query:
            sc = tsc = ln
        assert [int(x) for x in a[1:]] == [exp_qb, exp_qe, exp_rb, exp_re, sc, tsc, int(rec[i]["w"])]
