"""Shared test helpers: seed-task construction and comparisons (test infrastructure)."""
import numpy as np


def flat_from_lists(queries, targets):
    qoff = np.zeros(len(queries) + 1, dtype=np.int64)
    toff = np.zeros(len(targets) + 1, dtype=np.int64)
    np.cumsum([len(q) for q in queries], out=qoff[1:])
    np.cumsum([len(t) for t in targets], out=toff[1:])
    qbuf = np.concatenate([np.asarray(q, dtype=np.uint8) for q in queries] + [np.zeros(8, np.uint8)])
    tbuf = np.concatenate([np.asarray(t, dtype=np.uint8) for t in targets] + [np.zeros(8, np.uint8)])
    return qbuf, qoff, tbuf, toff


def seeds_from_flat(t, n_reads, unset_score_every=0):
    """Pair flat tasks (2r = reversed left flank, 2r+1 = right flank) into level-2 seed tasks."""
    seeds = []
    for r in range(n_reads):
        l, g = 2 * r, 2 * r + 1
        ql = t["qbuf"][t["qoff"][l]:t["qoff"][l + 1]]
        tl = t["tbuf"][t["toff"][l]:t["toff"][l + 1]]
        qr = t["qbuf"][t["qoff"][g]:t["qoff"][g + 1]]
        tr = t["tbuf"][t["toff"][g]:t["toff"][g + 1]]
        h0 = int(t["h0"][l])
        kind = r % 4
        if kind == 1:      # no left flank: BWA sets a->score = seed_len*a and skips the left extension
            ql, tl = ql[:0], tl[:0]
        elif kind == 2:    # no right flank
            qr, tr = qr[:0], tr[:0]
        init = h0 if (len(ql) == 0 or not unset_score_every or r % unset_score_every) else -1
        seeds.append(dict(q_left=ql, q_right=qr, t_left=tl, t_right=tr, init_score=init, qbeg=len(ql), h0=h0, id=1000 + r))
    return seeds


def assert_same(a, b, what=""):
    bad = np.nonzero(a != b)[0]
    assert len(bad) == 0, f"{what}: {len(bad)} mismatches, first at {bad[0]}: {a[bad[0]]} vs {b[bad[0]]}"


def oracle_params(O, p):
    """Byte-copy a bsw_b200.Params / Params2 into the oracle's own ctypes class of the same layout."""
    import ctypes as C
    cls = O.Params2 if hasattr(p, "pen_clip5") else O.Params
    q = cls()
    assert C.sizeof(q) == C.sizeof(p)
    C.memmove(C.byref(q), C.byref(p), C.sizeof(p))
    return q


def oracle_chain2aln(O, B, params2, seeds, variant=1):
    tasks, keep = B.make_seed_tasks(seeds)
    out, cells = O.chain2aln_batch(oracle_params(O, params2), tasks, variant=variant)
    del keep
    return out, cells


def random_small_tasks(rng, n, qmax=60, tmax=90):
    """Adversarial little tasks: every combination of tiny / small h0, w, lengths, related / unrelated / repetitive targets."""
    qs, ts, h0, w = [], [], [], []
    for k in range(n):
        ql = int(rng.integers(1, qmax + 1))
        tl = int(rng.integers(1, tmax + 1))
        mode = k % 5
        if mode == 0:
            q = rng.integers(0, 4, ql)
            t = rng.integers(0, 4, tl)
        elif mode == 1:                                  # low-complexity: many ties for the arg-max
            q = rng.integers(0, 2, ql)
            t = rng.integers(0, 2, tl)
        else:
            q = rng.integers(0, 4, ql)
            t = np.resize(q, tl)
            flips = rng.random(tl) < (0.03 if mode == 2 else 0.2)
            t = np.where(flips, (t + rng.integers(1, 4, tl)) % 4, t)
            if mode == 4 and tl > 6:                     # an indel
                c = int(rng.integers(1, tl - 2))
                t = np.concatenate([t[:c], t[c + int(rng.integers(1, 3)):], rng.integers(0, 4, 2)])[:tl]
        qs.append(q.astype(np.uint8)); ts.append(np.asarray(t).astype(np.uint8))
        h0.append(int(rng.choice([1, 2, 3, 5, 8, 13, 21, 40, 90])))
        w.append(int(rng.choice([0, 1, 2, 3, 5, 10, 30, 100])))
    qbuf, qoff, tbuf, toff = flat_from_lists(qs, ts)
    return dict(qbuf=qbuf, qoff=qoff, tbuf=tbuf, toff=toff, h0=np.array(h0, np.int32), w=np.array(w, np.int32))


def parse_tbb(tbb):
    """Field map: proc_element.v:815-820,915-918 (header), :880-892,871-874,826-828,924-934,807 (param words),
    task_parse.v:1924-1936 (data offsets), proc_element.v:1638,1677 (MS nibble first)."""
    hdr = dict(o_del=tbb[0] & 0xff, e_del=(tbb[0] >> 8) & 0xff, o_ins=(tbb[0] >> 16) & 0xff, e_ins=tbb[0] >> 24,
               pen_clip5=tbb[1] & 0xff, pen_clip3=(tbb[1] >> 8) & 0xff, w=(tbb[1] >> 16) & 0xff, n=int(tbb[2]))
    n = hdr["n"]
    off0 = int(tbb[8 + 2])
    tasks = []
    for i in range(n):
        pw = [int(x) for x in tbb[8 + 8 * i: 16 + 8 * i]]
        ql, tl = [pw[0] & 0xff, pw[1] & 0xff], [(pw[0] >> 16) & 0x7ff, (pw[1] >> 16) & 0x7ff]
        nb = ql[0] + ql[1] + tl[0] + tl[1]
        base = 8 + 8 * n + (pw[2] - off0)
        bases = [(int(tbb[base + (k >> 3)]) >> (28 - 4 * (k & 7))) & 15 for k in range(nb)]
        tasks.append(dict(qlen=ql, tlen=tl, init_score=pw[3] & 0xffff, qbeg=pw[3] >> 16, h0=pw[4] & 0xff,
                          max_ins=[pw[5] & 0xffff, pw[6] & 0xffff], max_del=[pw[5] >> 16, pw[6] >> 16], id=pw[7], bases=bases))
    return hdr, tasks
