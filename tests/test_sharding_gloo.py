"""
Multi-rank host logic on CPU (gloo, world size 2): every rank derives its own shard of the seeded
candidate stream (rank * n_per_rank), the shards tile the stream exactly, and the only
"collective" is the host-side gather of per-rank row counts -- the scoring path itself needs none.
"""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp
import torch

from pulsarfeatureextractor_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_per_rank, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = synth.generate(2, rank * n_per_rank, n_per_rank, want_subbands=False)
    digest = torch.tensor([int(b.profile.astype(np.int64).sum()), int(b.dmstat.astype(np.int64).sum()), len(b)],
                          dtype=torch.int64)
    gathered = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, digest)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put(([g.tolist() for g in gathered], float(t[0])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_tiles_the_stream():
    world, n_per_rank = 2, 3000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_per_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = synth.generate(2, 0, world * n_per_rank, want_subbands=False)
    for r in range(world):
        part = whole.slice(r * n_per_rank, (r + 1) * n_per_rank)
        assert gathered[r] == [int(part.profile.astype(np.int64).sum()), int(part.dmstat.astype(np.int64).sum()), n_per_rank]
    assert tmax == float(world)
