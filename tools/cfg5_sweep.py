"""BASELINE config 5: the 100 M-task whole-box sweep.  100 chunks of 1 M cfg2 tasks (regenerated from seed + chunk id, not
stored) are streamed through ONE context that owns N GPUs -- bsw_init(ids, N): the host deals the chunks of every call
to the devices, no collective (the FPGA analogue: batch_manager.v:397-562 dealing batches to its four arrays) -- for
N = 1, 2, 4, 8; a 1 % sample of every chunk is checked against the oracle.  One JSON line per N.
   python tools/cfg5_sweep.py [chunks=100] [gpus=1,2,4,8] > profiles/r02_cfg5_sweep.jsonl
Producer threads regenerate the next chunks while the GPUs work; the timed quantity is the time inside the batch calls
(generation is the caller's business), the wall time of the sweep is reported as well."""
import json, os, queue, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bsw_b200 as B
import oracle as O
import torch

nchunks = int(sys.argv[1]) if len(sys.argv) > 1 else 100
gpus = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8").split(",")]
M = 1_000_000
ndev = torch.cuda.device_count()
p, po = B.make_params(), O.make_params()
ncpu = os.cpu_count() or 8

QCAP, TCAP = 72_000_000, 136_000_000          # bytes of one 1 M-task chunk's query / target bases (cfg2: ~65 / ~125 MB)


def producer(q, pool, groups, per_call):
    """One batch call takes `per_call` consecutive chunks (one per GPU), generated side by side into one buffer pair."""
    for g in groups:
        qb, tb = pool.get()                      # a free pair of the caller's registered batch buffers
        # the generator indexes tasks globally, so per_call consecutive chunks are simply tasks [g*per_call*M, +per_call*M)
        t = B.synth_tasks("cfg2_150bp", per_call * M, first=g * per_call * M, qbuf=qb, tbuf=tb)
        q.put((g, t, (qb, tb)))


for N in gpus:
    if N > ndev:
        continue
    ctx = B.Context(devices=list(range(N)), host_threads=max(4, min(ncpu, 8 * N)))
    q, pool = queue.Queue(), queue.Queue()
    per_call = N                                 # chunks per batch call: one per GPU
    ngroups = nchunks // per_call
    for _ in range(4):                           # the host's ring of batch buffers, page-locked once
        qb, tb = np.zeros(QCAP * per_call, np.uint8), np.zeros(TCAP * per_call, np.uint8)
        ctx.register_host(qb); ctx.register_host(tb)
        pool.put((qb, tb))
    nprod = 3
    th = [threading.Thread(target=producer, args=(q, pool, range(i, ngroups, nprod), per_call), daemon=True) for i in range(nprod)]
    for x in th: x.start()
    # warm-up call (buffers, kernels): a host does this once when it opens the context
    wq, wt = np.zeros(QCAP, np.uint8), np.zeros(TCAP, np.uint8)
    t = B.synth_tasks("cfg2_150bp", M, first=nchunks * M, qbuf=wq, tbuf=wt)
    ctx.sw_extend_batch(p, t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    ctx.reset_stats()
    callers = 1                                  # one blocking call at a time, N chunks (one per GPU) per call
    lock = threading.Lock()
    acc = dict(in_calls=0.0, cells=0, ok=True, sampled=0, per_call=[])

    def caller(cid):
        try:
            caller_body(cid)
        except BaseException as e:               # never leave the producers blocked on the buffer ring
            acc["ok"] = False
            print("caller failed:", repr(e), file=sys.stderr, flush=True)
            os._exit(1)

    def caller_body(cid):
        out = np.zeros(M * per_call, dtype=B.RESULT_DTYPE)
        ctx.register_host(out)
        rng = np.random.default_rng(5 + cid)
        while True:
            item = q.get()
            if item is None:
                return
            k, t, bufs = item
            flat = (t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
            t0 = time.perf_counter()
            r, c = ctx.sw_extend_batch(p, *flat, out=out)
            dt = time.perf_counter() - t0
            # 1 % oracle sample: a random window of 10 000 consecutive tasks of every chunk of the call
            good = True
            for j in range(per_call):
                s0 = j * M + int(rng.integers(0, M - 10_000))
                sl = slice(s0, s0 + 10_000)
                ro, co = O.extend_batch(po, t["qbuf"], t["qoff"][s0:s0 + 10_001], t["tbuf"], t["toff"][s0:s0 + 10_001], t["h0"][sl], t["w"][sl], nthreads=4)
                good = good and bool(np.array_equal(ro, r[sl]) and np.array_equal(co, c[sl].astype(np.int64)))
            cells = int(c.astype(np.int64).sum())
            pool.put(bufs)
            with lock:
                acc["in_calls"] += dt; acc["cells"] += cells; acc["ok"] = acc["ok"] and good; acc["sampled"] += 10_000 * per_call; acc["per_call"].append(dt * 1e3)

    t_wall = time.perf_counter()
    cth = [threading.Thread(target=caller, args=(i,)) for i in range(callers)]
    for x in cth: x.start()
    for x in th: x.join()
    for _ in cth: q.put(None)
    for x in cth: x.join()
    wall = time.perf_counter() - t_wall
    st = ctx.stats()
    call_ms, cells_total = acc["per_call"], acc["cells"]
    print(json.dumps(dict(config="cfg5 sweep: %d x 1M cfg2 tasks through ONE context, bsw_init(ids, %d), %d chunk(s) per batch call" % (ngroups * per_call, N, per_call),
                          n_gpus=N, tasks=ngroups * per_call * M, cells=cells_total, sweep_wall_s=wall, gcups_e2e=cells_total / wall * 1e-9,
                          mtasks_per_s_e2e=ngroups * per_call * M / wall * 1e-6, ms_per_call_median=float(np.median(call_ms)),
                          gcups_inside_calls=cells_total / (acc["in_calls"] / callers) * 1e-9,
                          mtasks_per_s_inside_calls=ngroups * per_call * M / (acc["in_calls"] / callers) * 1e-6,
                          oracle_sampled_tasks=acc["sampled"], bit_exact_sample=acc["ok"],
                          h2d_bytes_per_chunk=st["h2d_bytes"] // (ngroups * per_call), host_cores=ncpu,
                          note="wall time of the whole sweep (chunk generation by 4 producer threads and the oracle sample run beside the batch calls)")), flush=True)
    ctx.close()
