"""Seeded synthetic inputs / weights shared by the golden generator, the tests and bench.py.
TEST INFRASTRUCTURE (see oracle/__init__.py).  Distributions follow SURVEY.md 8(d)."""
from __future__ import annotations

from typing import Dict

import torch

Tensor = torch.Tensor


def encoder_init_state(channels: int, seed: int = 13) -> Dict[str, Tensor]:
    """Default ``nn.Module`` init of the reference Encoder under ``seed`` WITHOUT the reference tree:
    same module construction order as /root/reference/model.py:43-57,90-101 (Conv1d, LayerNorm,
    [Linear, LayerNorm]x4, Linear, codebook buffers, LSTM), hence the same RNG stream.  Verified equal
    to the live reference by tests/test_oracle_cpu.py::test_init_matches_reference."""
    import torch.nn as nn

    torch.manual_seed(seed)
    C = channels
    sd: Dict[str, Tensor] = {}
    conv = nn.Conv1d(80, C, 4, 2, 1, bias=False)
    sd["conv.weight"] = conv.weight.detach().clone()
    ln0 = nn.LayerNorm(C)
    sd["encoder.0.weight"], sd["encoder.0.bias"] = ln0.weight.detach().clone(), ln0.bias.detach().clone()
    for j in range(4):
        lin = nn.Linear(C, C, bias=False)
        ln = nn.LayerNorm(C)
        sd[f"encoder.{2 + 3 * j}.weight"] = lin.weight.detach().clone()
        sd[f"encoder.{3 + 3 * j}.weight"] = ln.weight.detach().clone()
        sd[f"encoder.{3 + 3 * j}.bias"] = ln.bias.detach().clone()
    proj = nn.Linear(C, 64)
    sd["encoder.14.weight"], sd["encoder.14.bias"] = proj.weight.detach().clone(), proj.bias.detach().clone()
    emb = torch.Tensor(512, 64)
    emb.uniform_(-1 / 512, 1 / 512)
    sd["codebook.embedding"] = emb
    sd["codebook.ema_count"] = torch.zeros(512)
    sd["codebook.ema_weight"] = emb.clone()
    rnn = nn.LSTM(64, 256, batch_first=True)
    for k, v in rnn.state_dict().items():
        sd[f"rnn.{k}"] = v.detach().clone()
    return sd


def perturb_encoder_state(sd: Dict[str, Tensor], seed: int = 99) -> Dict[str, Tensor]:
    """"Trained-like" variant: non-trivial LayerNorm affine and an N(0, 0.5) codebook, so the affine
    terms and a well-separated codebook are exercised (default init has LN weight 1 / bias 0)."""
    g = torch.Generator().manual_seed(seed)
    out = {k: v.clone() for k, v in sd.items()}
    for i in (0, 3, 6, 9, 12):
        out[f"encoder.{i}.weight"] = out[f"encoder.{i}.weight"] + 0.1 * torch.randn(out[f"encoder.{i}.weight"].shape, generator=g)
        out[f"encoder.{i}.bias"] = out[f"encoder.{i}.bias"] + 0.1 * torch.randn(out[f"encoder.{i}.bias"].shape, generator=g)
    out["codebook.embedding"] = 0.5 * torch.randn(512, 64, generator=g)
    out["codebook.ema_weight"] = out["codebook.embedding"].clone()
    return out


def synthetic_mel(B: int, T: int, seed: int = 0, kind: str = "rand") -> Tensor:
    """80-bin log-mel-like input: uniform [0,1) (the reference normalises log-mel to ~[0,1],
    /root/reference/preprocess.py:73-74) or the randn stress set."""
    g = torch.Generator().manual_seed(seed)
    if kind == "rand":
        return torch.rand(B, 80, T, generator=g)
    return torch.randn(B, 80, T, generator=g)


def vq_inputs(n_frames: int, kind: str = "init", seed: int = 1234, batch: int = 1):
    """VQ microbench inputs.  ``init``: codebook U(+-1/512), x ~ N(0,1) (worst case for
    ties).  ``trained``: e ~ N(0,1), x = e[randint] + 0.3 N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    if kind == "init":
        cb = (torch.rand(512, 64, generator=g) * 2 - 1) / 512
        x = torch.randn(batch, n_frames, 64, generator=g)
    else:
        cb = torch.randn(512, 64, generator=g)
        pick = torch.randint(0, 512, (batch, n_frames), generator=g)
        x = cb[pick] + 0.3 * torch.randn(batch, n_frames, 64, generator=g)
    return x, cb


def vocoder_inputs(B: int, Tc: int, seed: int = 0, n_speakers: int = 102, n_steps: int | None = None):
    """codes (B,Tc) in [0,512), speaker (B,) in [0,n_speakers) under ``seed``; uniforms (B,L) under seed 7."""
    g = torch.Generator().manual_seed(seed)
    codes = torch.randint(0, 512, (B, Tc), generator=g)
    speaker = torch.randint(0, n_speakers, (B,), generator=g)
    L = 320 * Tc if n_steps is None else n_steps
    gu = torch.Generator().manual_seed(7)
    uniforms = torch.rand(B, L, generator=gu)
    return codes, speaker, uniforms


def state_checksum(sd: Dict[str, Tensor]) -> float:
    """Order-independent fingerprint of a weight set (float64 sum of |w|)."""
    return float(sum(v.double().abs().sum() for v in sd.values()))
