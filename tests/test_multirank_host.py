"""N>1 path on the CPU: two gloo ranks shard one workload exactly as bench.py does (rank r owns tasks
[r*n, (r+1)*n), no data-path collective) and run the host logic + K1 lane emulation; the per-rank checksums gathered
over torch.distributed must equal a single-process run over the union."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch
    import torch.distributed as dist
    import bsw_b200 as B
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t = B.synth_tasks("cfg3_mixed", n, first=rank * n, seed=5)
    res, cells, _ = B.emu_extend_batch(B.make_params(), t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    mine = torch.tensor([int(res["score"].astype(np.int64).sum()), int(res["qle"].astype(np.int64).sum()),
                         int(res["max_off"].astype(np.int64).sum()), int(cells.astype(np.int64).sum()), n], dtype=torch.int64)
    dist.barrier()
    tot = mine.clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    tmax = torch.tensor([float(rank + 1)], dtype=torch.float64)          # the max-over-ranks timing plumbing of bench.py
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put((tot.tolist(), float(tmax[0])))
    dist.destroy_process_group()


def test_two_ranks_shard_without_collective(B):
    import torch.multiprocessing as mp
    n, world = 1500, 2
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    port = _free_port()
    procs = [ctxmp.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    tot, tmax = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    t = B.synth_tasks("cfg3_mixed", world * n, first=0, seed=5)
    res, cells, _ = B.emu_extend_batch(B.make_params(), t["qbuf"], t["qoff"], t["tbuf"], t["toff"], t["h0"], t["w"])
    want = [int(res["score"].astype(np.int64).sum()), int(res["qle"].astype(np.int64).sum()),
            int(res["max_off"].astype(np.int64).sum()), int(cells.astype(np.int64).sum()), world * n]
    assert tot == want and tmax == float(world)


def test_shards_regenerate_independently(B):
    """Config 5 streams chunks regenerated from (seed, first): a chunk equals the same range of a bigger draw."""
    a = B.synth_tasks("cfg2_150bp", 3000, first=0, seed=2)
    b = B.synth_tasks("cfg2_150bp", 1000, first=2000, seed=2)
    lo = a["qoff"][2000]
    assert np.array_equal(a["qbuf"][lo:a["qoff"][3000]], b["qbuf"][:b["qoff"][1000]])
    assert np.array_equal(a["h0"][2000:], b["h0"])
