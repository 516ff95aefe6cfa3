"""CPU restatement of ``Vocoder.generate`` / ``Vocoder.forward`` -- TEST INFRASTRUCTURE.

**PARITY UNPINNED** (see oracle/__init__.py): the in-tree part -- code/speaker embedding, x2 nearest
upsample, broadcast, concat (/root/reference/network_vocoder.py:41-78) -- is restated exactly; the
``rnnms`` part (prenet biGRU, x160 upsample, autoregressive loop) is NOT in the reference tree and is
restated from the published RNN_MS/WaveRNN algorithm with the dimensions pinned by
/root/reference/config.py:62-77,199.  PyTorch ``nn.GRU`` / ``nn.GRUCell`` gate
conventions (rows ordered r, z, n):

    r = s(W_ir x + b_ir + W_hr h + b_hr) ; z = s(W_iz x + b_iz + W_hz h + b_hz)
    n = tanh(W_in x + b_in + r * (W_hn h + b_hn)) ; h' = (1 - z) * n + z * h

Sampling: ``Categorical(softmax(o)).sample()`` cannot take injected randomness, so the oracle and the
CUDA kernel share this definition (inverse CDF over the fp32 softmax with one injected uniform u_t per
(utterance, step)):

    m = max_k o_k ; e_k = exp(o_k - m) ; c_k = sum_{j<=k} e_j ; S = c_255
    x_t = min{ k : c_k > u_t * S }      (255 if no k qualifies, i.e. u_t*S rounds up to S)

Canonical state_dict keys (names under ``rnnms.`` are OUR choice -- upstream names unknown; the product
loader carries a remap table):
    code_embedding.weight (512,64) ; speaker_embedding.weight (n_spk,64)
    rnnms.prenet.net.{weight_ih,weight_hh,bias_ih,bias_hh}_l{0,1}[_reverse]   biGRU(128 -> 128/dir, 2 layers)
    rnnms.ar.embedding.weight (256,256)
    rnnms.ar.rnn.{weight_ih_l0 (2688,512), weight_hh_l0 (2688,896), bias_ih_l0, bias_hh_l0}
    rnnms.ar.fc1.{weight (256,896), bias} ; rnnms.ar.fc2.{weight (256,256), bias}
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .mulaw import mulaw_decode_lut

Tensor = torch.Tensor
UPSAMPLE_T = 160       # config.py:70,102  (hop_length)
X_INIT = 128           # 2**bits / 2, first AR input


def init_state_dict(n_speakers: int = 102, seed: int = 13, size_i_codebook: int = 512, dim_i_embedding: int = 64,
                    dim_speaker_embedding: int = 64, dim_voc_latent: int = 256, size_i_embed_ar: int = 256,
                    size_h_rnn: int = 896, size_h_fc: int = 256, bits_mu_law: int = 8) -> Dict[str, Tensor]:
    """Random-init weights with default ``nn.Module`` init under ``seed`` (config.py:13), built from the
    dims in config.py:62-77,199.  Module creation order: code_embedding, speaker_embedding
    (network_vocoder.py:37-38), then prenet GRU, AR embedding, AR GRU, fc1, fc2."""
    import torch.nn as nn

    torch.manual_seed(seed)
    dim_i_feature = dim_i_embedding + dim_speaker_embedding     # config.py:199
    code_emb = nn.Embedding(size_i_codebook, dim_i_embedding)
    spk_emb = nn.Embedding(n_speakers, dim_speaker_embedding)
    prenet = nn.GRU(dim_i_feature, dim_voc_latent // 2, num_layers=2, batch_first=True, bidirectional=True)
    emb = nn.Embedding(1 << bits_mu_law, size_i_embed_ar)
    rnn = nn.GRU(size_i_embed_ar + dim_voc_latent, size_h_rnn, batch_first=True)
    fc1 = nn.Linear(size_h_rnn, size_h_fc)
    fc2 = nn.Linear(size_h_fc, 1 << bits_mu_law)
    sd: Dict[str, Tensor] = {}
    sd["code_embedding.weight"] = code_emb.weight.detach().clone()
    sd["speaker_embedding.weight"] = spk_emb.weight.detach().clone()
    for k, v in prenet.state_dict().items():
        sd[f"rnnms.prenet.net.{k}"] = v.detach().clone()
    sd["rnnms.ar.embedding.weight"] = emb.weight.detach().clone()
    for k, v in rnn.state_dict().items():
        sd[f"rnnms.ar.rnn.{k}"] = v.detach().clone()
    sd["rnnms.ar.fc1.weight"] = fc1.weight.detach().clone()
    sd["rnnms.ar.fc1.bias"] = fc1.bias.detach().clone()
    sd["rnnms.ar.fc2.weight"] = fc2.weight.detach().clone()
    sd["rnnms.ar.fc2.bias"] = fc2.bias.detach().clone()
    return sd


def embed_inputs(sd: Dict[str, Tensor], z: Tensor, speaker: Tensor) -> Tensor:
    """/root/reference/network_vocoder.py:73-77: code embedding, nearest x2 (frame t>>1), speaker embedding
    broadcast over time, concat.  z (B,Tc) int64, speaker (B,) int64 -> (B, 2Tc, 128)."""
    zc = sd["code_embedding.weight"][z]                          # (B,Tc,64)
    zc2 = zc.repeat_interleave(2, dim=1)                         # F.interpolate(scale_factor=2), nearest
    s = sd["speaker_embedding.weight"][speaker]                  # (B,64)
    s2 = s[:, None, :].expand(-1, zc2.shape[1], -1)
    return torch.cat((zc2, s2), dim=-1)


def gru_cell(x_proj: Tensor, h: Tensor, w_hh: Tensor, b_hh: Tensor) -> Tensor:
    """One GRU step given a = W_ih x + b_ih (x_proj).  Rows r,z,n."""
    H = h.shape[-1]
    b = h @ w_hh.t() + b_hh
    r = torch.sigmoid(x_proj[..., :H] + b[..., :H])
    zg = torch.sigmoid(x_proj[..., H:2 * H] + b[..., H:2 * H])
    n = torch.tanh(x_proj[..., 2 * H:] + r * b[..., 2 * H:])
    return (1.0 - zg) * n + zg * h


def gru_direction(u: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor, reverse: bool) -> Tensor:
    B, T, _ = u.shape
    H = w_hh.shape[1]
    xp = u @ w_ih.t() + b_ih
    h = torch.zeros(B, H, dtype=u.dtype)
    out = torch.empty(B, T, H, dtype=u.dtype)
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        h = gru_cell(xp[:, t], h, w_hh, b_hh)
        out[:, t] = h
    return out


def prenet(sd: Dict[str, Tensor], u: Tensor, num_layers: int = 2) -> Tensor:
    """2-layer bidirectional GRU, hidden 128/direction (config.py:68,71-73); layer-1 input = [fwd;bwd] of
    layer 0.  (B,2Tc,128) -> (B,2Tc,256)."""
    x = u
    for layer in range(num_layers):
        outs = []
        for suffix, rev in (("", False), ("_reverse", True)):
            p = f"rnnms.prenet.net.{{}}_l{layer}{suffix}"
            outs.append(gru_direction(x, sd[p.format("weight_ih")], sd[p.format("weight_hh")],
                                      sd[p.format("bias_ih")], sd[p.format("bias_hh")], rev))
        x = torch.cat(outs, dim=-1)
    return x


def condition(sd: Dict[str, Tensor], z: Tensor, speaker: Tensor) -> Tensor:
    """Per-frame conditioning p (B,2Tc,256); the x160 nearest upsample is the index map t -> t // 160."""
    return prenet(sd, embed_inputs(sd, z, speaker))


def ar_logits_step(sd: Dict[str, Tensor], x_prev: Tensor, cond_t: Tensor, h: Tensor) -> Tuple[Tensor, Tensor]:
    """One AR step AS THE REFERENCE WRITES IT (unhoisted): embedding -> cat -> GRUCell -> fc1 -> ReLU -> fc2."""
    e = sd["rnnms.ar.embedding.weight"][x_prev]                         # (B,256)
    inp = torch.cat((e, cond_t), dim=-1)                                # (B,512)
    a = inp @ sd["rnnms.ar.rnn.weight_ih_l0"].t() + sd["rnnms.ar.rnn.bias_ih_l0"]
    h = gru_cell(a, h, sd["rnnms.ar.rnn.weight_hh_l0"], sd["rnnms.ar.rnn.bias_hh_l0"])
    y = torch.relu(h @ sd["rnnms.ar.fc1.weight"].t() + sd["rnnms.ar.fc1.bias"])
    o = y @ sd["rnnms.ar.fc2.weight"].t() + sd["rnnms.ar.fc2.bias"]
    return o, h


def sample_inverse_cdf(o: Tensor, u: Tensor) -> Tensor:
    """x = min{k : cumsum(exp(o - max))_k > u * S} (see module docstring).  o (B,256), u (B,) -> (B,) int64."""
    m = o.max(dim=-1, keepdim=True).values
    e = torch.exp(o - m)
    c = torch.cumsum(e, dim=-1)
    thr = (u * c[:, -1])[:, None]
    hit = c > thr
    first = torch.where(hit.any(dim=-1), hit.to(torch.int64).argmax(dim=-1),
                        torch.full((o.shape[0],), o.shape[-1] - 1, dtype=torch.int64))
    return first


def generate(sd: Dict[str, Tensor], z: Tensor, speaker: Tensor, uniforms: Tensor,
             n_steps: Optional[int] = None, return_all: bool = False):
    """``Vocoder.generate(z, speaker)`` (/root/reference/network_vocoder.py:69-78 + rnnms restated).

    uniforms (B, L) in [0,1).  Returns wav (B,L) float32 = mu-law decode of the sampled codes; with
    ``return_all`` also (codes (B,L) int64, logits (B,L,256)).
    """
    p = condition(sd, z, speaker)
    B, T2, _ = p.shape
    L = T2 * UPSAMPLE_T if n_steps is None else n_steps
    H = sd["rnnms.ar.rnn.weight_hh_l0"].shape[1]
    h = torch.zeros(B, H, dtype=p.dtype)
    x = torch.full((B,), X_INIT, dtype=torch.int64)
    codes = torch.empty(B, L, dtype=torch.int64)
    logits = torch.empty(B, L, 256, dtype=p.dtype) if return_all else None
    for t in range(L):
        o, h = ar_logits_step(sd, x, p[:, t // UPSAMPLE_T], h)
        x = sample_inverse_cdf(o.float(), uniforms[:, t].float())
        codes[:, t] = x
        if return_all:
            logits[:, t] = o
    lut = torch.from_numpy(mulaw_decode_lut(8))
    wav = lut[codes]
    if return_all:
        return wav, codes, logits
    return wav


def forward_teacher_forced(sd: Dict[str, Tensor], x: Tensor, z: Tensor, speaker: Tensor) -> Tensor:
    """``Vocoder.forward(x, z, speaker)`` (/root/reference/network_vocoder.py:41-67; I/O contract
    /root/reference/vocoder.py:62-63): x (B,L) int64 mu-law series (the AR input at step t is x[:,t]),
    returns un-normalised energies (B,L,256)."""
    p = condition(sd, z, speaker)
    B, L = x.shape
    H = sd["rnnms.ar.rnn.weight_hh_l0"].shape[1]
    h = torch.zeros(B, H, dtype=p.dtype)
    out = torch.empty(B, L, 256, dtype=p.dtype)
    for t in range(L):
        o, h = ar_logits_step(sd, x[:, t], p[:, t // UPSAMPLE_T], h)
        out[:, t] = o
    return out


def cdf_bounds(logits: Tensor) -> Tensor:
    """fp64 normalised inclusive CDF (…,256) of softmax(logits): used to check that a sampled code is
    CONSISTENT with its uniform (c_{x-1} - tol <= u <= c_x + tol) without requiring bit-equal rounding."""
    pr = torch.softmax(logits.double(), dim=-1)
    return torch.cumsum(pr, dim=-1)


def to_numpy_state(sd: Dict[str, Tensor]) -> Dict[str, np.ndarray]:
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}
