"""world_size-2 gloo test of the multi-rank host path: shot sharding + the one collective (counter all-reduce)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shots, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qldpcsim_b200 import simulator
    assert simulator.dist_info() == (rank, world)
    lo, hi = simulator.shard_range(shots, rank, world)
    # stand-in for the per-rank device counters: classify a deterministic synthetic outcome per global shot
    s = np.arange(lo, hi)
    c = torch.tensor([(s % 7 == 0).sum(), (s % 11 == 0).sum(), (s % 3 != 0).sum(), 0, (s % 50 + 1).sum(), (s % 49 + 1).sum(),
                      hi - lo, 0], dtype=torch.int64)
    simulator.reduce_counters(c)
    q.put((rank, c.tolist()))
    dist.destroy_process_group()


def test_two_rank_counter_reduction():
    world, shots = 2, 100003
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shots, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s = np.arange(shots)
    want = [int((s % 7 == 0).sum()), int((s % 11 == 0).sum()), int((s % 3 != 0).sum()), 0, int((s % 50 + 1).sum()),
            int((s % 49 + 1).sum()), shots, 0]
    assert out[0] == want and out[1] == want          # identical on every rank, independent of the sharding
