"""CPU, world_size 2, gloo: the sharding / gather host logic of the multi-GPU path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vectorquantizedcpc_b200 import dist as vdist


def test_shard_ranges_partition_everything():
    for n in (0, 1, 2, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                lo, hi = vdist.shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                covered += list(range(lo, hi))
            assert covered == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n * 6, dtype=torch.float32).reshape(n, 6)
        idx_full = torch.arange(n * 3, dtype=torch.int64).reshape(n, 3)
        local = vdist.shard(full) * 2.0          # stand-in for the per-rank compute
        out = vdist.gather_utterances(local, n, dst=0)
        out_i = vdist.gather_utterances(vdist.shard(idx_full), n, dst=0)
        if rank == 0:
            q.put((torch.equal(out, full * 2.0), torch.equal(out_i, idx_full), tuple(out.shape)))
        else:
            assert out is None and out_i is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [5, 8, 1])
def test_gather_utterances_world2_gloo(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    ok, ok_i, shape = q.get(timeout=5)
    assert ok and ok_i and shape == (n, 6)
