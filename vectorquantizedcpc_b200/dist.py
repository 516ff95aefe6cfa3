"""Utterance-level data parallelism: utterances are independent in every stage of the path
(/root/reference/model.py:59-70, network_vocoder.py:69-78), so ranks take contiguous blocks of utterances,
weights are replicated and the ONLY collective is one gather of the outputs.  One process per GPU
(torchrun); backend nccl on GPUs, gloo for the CPU tests of this host logic."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of ``n`` utterances for ``rank``: ceil(n/world) per rank, trailing ranks may be short/empty."""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def shard(t: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    """This rank's block of utterances (dim 0)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_range(t.shape[0], rank, world)
    return t[lo:hi]


def gather_utterances(local: torch.Tensor, n_total: int, dst: int = 0) -> Optional[torch.Tensor]:
    """Gather per-rank blocks (dim 0, possibly ragged / empty) back into utterance order on ``dst``: ONE fixed-size
    ``gather`` straight into the slices of the result -- with ceil(n/world) utterances per rank only trailing ranks are
    short or empty, so the valid rows form a prefix of the (world * per) buffer and nothing is copied afterwards; a full
    block is sent as it is, a short one padded to the common size.  Returns the full tensor on ``dst`` and None elsewhere."""
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (n_total + world - 1) // world
    local = local.contiguous()
    if local.shape[0] == per:
        send = local
    else:
        send = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        send[: local.shape[0]] = local
    out = None
    bufs: Optional[List[torch.Tensor]] = None
    if rank == dst:
        out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        bufs = [out[r * per:(r + 1) * per] for r in range(world)]
    dist.gather(send, bufs, dst=dst)
    return out[:n_total] if rank == dst else None


def convert_sharded(encoder, vocoder, mel: torch.Tensor, speaker: torch.Tensor, dst: int = 0, **generate_kw):
    """End-to-end conversion of a batch of equal-length utterances (convert.py:72-77) sharded over the ranks:
    every rank encodes + generates its block; indices and waveforms are gathered on ``dst``.

    ``generate_kw`` is passed to ``Vocoder.generate``; keywords that change its return type (``return_mulaw``,
    ``return_logits``) are rejected on EVERY rank before any collective starts (a rank raising inside the gather would
    leave the others blocked in it).  A per-utterance ``uniforms`` tensor (n, L) is sharded like the inputs."""
    bad = [k for k in ("return_mulaw", "return_logits") if generate_kw.get(k)]
    if bad:
        raise ValueError(f"convert_sharded gathers waveforms only; unsupported generate keywords: {bad}")
    n = mel.shape[0]
    m, s = shard(mel), shard(speaker)
    kw = dict(generate_kw)
    if kw.get("uniforms") is not None:
        kw["uniforms"] = shard(kw["uniforms"])
    if kw.get("lengths") is not None:
        kw["lengths"] = shard(torch.as_tensor(kw["lengths"]))
    Tp = (mel.shape[2] - 2) // 2 + 1
    up = 2 * int(vocoder.conf.rnnms.upsampling_t)        # samples per code frame: x2 nearest, then the hop (config.py:70,102)
    L = int(kw["n_steps"]) if kw.get("n_steps") is not None else up * Tp
    if m.shape[0] > 0:
        _, _, idx = encoder.encode(m)
        wav = vocoder.generate(idx, s, **kw)
    else:                                                 # an empty block still takes part in the gathers, with the same shapes
        idx = torch.empty(0, Tp, dtype=torch.int64, device=mel.device)
        wav = torch.empty(0, L, device=mel.device)
    return gather_utterances(idx, n, dst), gather_utterances(wav, n, dst)
