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


class _StubEncoder:
    """encode(mel (B, 80, T)) -> (z, c, idx) with idx a deterministic function of the utterance (CPU stand-in)."""

    def encode(self, mel):
        Tp = (mel.shape[2] - 2) // 2 + 1
        idx = (mel[:, 0, :Tp] * 100).round().long() % 512
        return None, None, idx


class _StubVocoder:
    class _Conf:
        class rnnms:
            upsampling_t = 160
    conf = _Conf()

    def generate(self, idx, speaker, n_steps=None, uniforms=None, **kw):
        L = 320 * idx.shape[1] if n_steps is None else n_steps
        base = idx[:, :1].float() + speaker[:, None].float()
        wav = base.expand(-1, L).clone()
        if uniforms is not None:
            assert uniforms.shape == (idx.shape[0], L)
            wav = wav + uniforms
        return wav


def _convert_worker(rank, world, port, n, n_steps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        mel = torch.rand(n, 80, 12, generator=g)
        spk = torch.arange(n)
        L = 320 * 6 if n_steps is None else n_steps
        u = torch.rand(n, L, generator=g)
        # keywords that change generate's return type are refused on every rank, before any collective
        try:
            vdist.convert_sharded(_StubEncoder(), _StubVocoder(), mel, spk, return_mulaw=True)
            refused = False
        except ValueError:
            refused = True
        idx, wav = vdist.convert_sharded(_StubEncoder(), _StubVocoder(), mel, spk, n_steps=n_steps, uniforms=u)
        if rank == 0:
            _, _, idx_ref = _StubEncoder().encode(mel)
            wav_ref = _StubVocoder().generate(idx_ref, spk, n_steps=n_steps, uniforms=u)
            q.put((refused, torch.equal(idx, idx_ref), torch.equal(wav, wav_ref), tuple(wav.shape)))
        else:
            assert idx is None and wav is None and refused
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,n_steps", [(1, 100), (3, None), (4, 777)])
def test_convert_sharded_world2_gloo_empty_shard_and_n_steps(n, n_steps):
    """n = 1 leaves rank 1 with an EMPTY block: its placeholder tensors must have the shapes the other rank gathers
    (n_steps, not 320 * T'), or the collective mismatches."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_convert_worker, args=(r, 2, port, n, n_steps, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    refused, ok_i, ok_w, shape = q.get(timeout=5)
    assert refused and ok_i and ok_w and shape == (n, 1920 if n_steps is None else n_steps)
