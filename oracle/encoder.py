"""CPU restatement of ``Encoder.encode`` -- TEST INFRASTRUCTURE (see oracle/__init__.py).

PINNED: every function here is checked against the live reference
(``tests/test_oracle_cpu.py`` + ``tests/golden/encoder_*.npz``).

Each function states the formula it follows and the reference line it restates.
All functions are written from ``state_dict`` tensors only (no ``nn.Module``
forward calls), in the dtype of their inputs, so that the same code gives the
fp32 restatement and -- after ``.double()`` -- the fp64 "truth" used to classify
near-ties.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

Tensor = torch.Tensor

LN_EPS = 1e-5  # nn.LayerNorm default, /root/reference/model.py:47,51


def conv_out_len(T: int) -> int:
    """T' of ``nn.Conv1d(80, C, 4, 2, 1)`` -- /root/reference/model.py:43."""
    return (T + 2 * 1 - 4) // 2 + 1


def conv_frontend(mel: Tensor, weight: Tensor) -> Tensor:
    """``self.conv(mel).transpose(1, 2)`` -- /root/reference/model.py:43,65,67.

    y[b,t,o] = sum_{i<80,k<4} W[o,i,k] * mel[b,i,2t+k-1]   (zero outside [0,T)), no bias.
    mel (B,80,T) -> (B,T',C)
    """
    B, I, T = mel.shape
    C = weight.shape[0]
    Tp = conv_out_len(T)
    padded = torch.zeros(B, I, T + 2, dtype=mel.dtype)
    padded[:, :, 1:T + 1] = mel
    t_idx = 2 * torch.arange(Tp)[:, None] + torch.arange(4)[None, :]        # (T',4) -> padded index 2t+k
    A = padded[:, :, t_idx]                                                  # (B,I,T',4)
    A = A.permute(0, 2, 1, 3).reshape(B, Tp, I * 4)                          # K index = i*4+k
    return A @ weight.reshape(C, I * 4).t()


def layer_norm(v: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """``nn.LayerNorm(C)``: biased variance, eps 1e-5, affine -- /root/reference/model.py:47,51."""
    mean = v.mean(dim=-1, keepdim=True)
    var = ((v - mean) ** 2).mean(dim=-1, keepdim=True)
    return (v - mean) / torch.sqrt(var + LN_EPS) * w + b


def fc_stack(y0: Tensor, sd: Dict[str, Tensor], return_hidden: bool = False):
    """``self.encoder`` = LN-ReLU-[Linear-LN-ReLU]x4-Linear(C,64) -- /root/reference/model.py:46-55,67.

    state_dict indices: LN at 0,3,6,9,12; Linear(no bias) at 2,5,8,11; projection (with bias) at 14.
    """
    y = torch.relu(layer_norm(y0, sd["encoder.0.weight"], sd["encoder.0.bias"]))
    for j in range(4):
        lin = sd[f"encoder.{2 + 3 * j}.weight"]
        y = y @ lin.t()
        y = torch.relu(layer_norm(y, sd[f"encoder.{3 + 3 * j}.weight"], sd[f"encoder.{3 + 3 * j}.bias"]))
    z_pre = y @ sd["encoder.14.weight"].t() + sd["encoder.14.bias"]
    if return_hidden:
        return z_pre, y
    return z_pre


def vq_distances(x_flat: Tensor, codebook: Tensor) -> Tensor:
    """d[n,m] = (|e_m|^2 + |x_n|^2) - 2 x_n.e_m  -- /root/reference/model.py:107-110 (addmm, alpha=-2, beta=1)."""
    e2 = (codebook ** 2).sum(dim=1)
    x2 = (x_flat ** 2).sum(dim=1, keepdim=True)
    return (e2 + x2) - 2.0 * (x_flat @ codebook.t())


def vq_lookup(x: Tensor, codebook: Tensor) -> Tuple[Tensor, Tensor]:
    """``VQEmbeddingEMA.encode`` -- /root/reference/model.py:103-115.

    x (B,T,D) -> quantized (B,T,D) = codebook[idx], idx (B,T) int64 = FIRST argmin of the distances
    (torch.argmin returns the lowest index on exact ties).
    """
    D = codebook.shape[1]
    flat = x.reshape(-1, D)
    d = vq_distances(flat, codebook)
    idx = torch.argmin(d.float(), dim=-1)
    q = codebook[idx].view_as(x)
    return q, idx.view(x.shape[0], x.shape[1])


def vq_scores_exact(x: Tensor, codebook: Tensor) -> Tensor:
    """fp64 ``|e|^2 - 2 x.e`` (the |x|^2 term is argmin-invariant).  Used only to CLASSIFY mismatches
    as near-ties: returns (N, M) float64."""
    D = codebook.shape[1]
    flat = x.reshape(-1, D).double()
    e = codebook.double()
    return (e ** 2).sum(dim=1)[None, :] - 2.0 * (flat @ e.t())


def lstm(z: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor) -> Tensor:
    """``nn.LSTM(64, 256, batch_first=True)`` output sequence -- /root/reference/model.py:57,69.

    g = W_ih z_t + b_ih + W_hh h + b_hh ; (i,f,g,o) = chunk4 ; c = s(f) c + s(i) tanh(g) ; h = s(o) tanh(c)
    h0 = c0 = 0.  z (B,T,64) -> (B,T,256)
    """
    B, T, _ = z.shape
    H = w_hh.shape[1]
    h = torch.zeros(B, H, dtype=z.dtype)
    c = torch.zeros(B, H, dtype=z.dtype)
    xp = z @ w_ih.t() + b_ih
    out = torch.empty(B, T, H, dtype=z.dtype)
    for t in range(T):
        g = xp[:, t] + h @ w_hh.t() + b_hh
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[:, t] = h
    return out


def encode(sd: Dict[str, Tensor], mel: Tensor, return_aux: bool = False):
    """``Encoder.encode(mel) -> (z, c, indices)`` -- /root/reference/model.py:59-70.

    ``z`` is the QUANTISED vector series (model.py:68,70).  With ``return_aux`` also returns the
    pre-VQ projection (what the forward hook on ``encoder.encoder[-1]`` sees, encode.py:34-40).
    """
    y0 = conv_frontend(mel, sd["conv.weight"])
    z_pre = fc_stack(y0, sd)
    z_q, idx = vq_lookup(z_pre, sd["codebook.embedding"])
    c = lstm(z_q, sd["rnn.weight_ih_l0"], sd["rnn.weight_hh_l0"], sd["rnn.bias_ih_l0"], sd["rnn.bias_hh_l0"])
    if return_aux:
        return z_q, c, idx, z_pre
    return z_q, c, idx


def classify_index_mismatches(x: Tensor, codebook: Tensor, idx_a: Tensor, idx_b: Tensor, slack: float = 0.0):
    """Compare two index maps for the same ``x``.  A mismatch is a NEAR-TIE iff the fp64 score gap between
    the two candidates is <= 4*ulp32(|x|^2 + |e|^2) + slack (rule adopted in SURVEY.md 7.2: the reference
    adds |x|^2 before the argmin, which quantises distances to that ulp).

    Returns dict(n, mismatches, near_ties, hard, max_gap).
    """
    D = codebook.shape[1]
    flat = x.reshape(-1, D)
    a = idx_a.reshape(-1)
    b = idx_b.reshape(-1)
    bad = torch.nonzero(a != b).flatten()
    out = dict(n=int(a.numel()), mismatches=int(bad.numel()), near_ties=0, hard=0, max_gap=0.0)
    if bad.numel() == 0:
        return out
    xs = flat[bad].double()
    e = codebook.double()
    sa = (e[a[bad]] ** 2).sum(1) - 2.0 * (xs * e[a[bad]]).sum(1)
    sb = (e[b[bad]] ** 2).sum(1) - 2.0 * (xs * e[b[bad]]).sum(1)
    gap = (sa - sb).abs()
    mag = ((xs ** 2).sum(1) + torch.maximum((e[a[bad]] ** 2).sum(1), (e[b[bad]] ** 2).sum(1))).float()
    ulp = torch.nextafter(mag, torch.full_like(mag, float("inf"))) - mag
    tol = 4.0 * ulp.double() + slack
    near = gap <= tol
    out["near_ties"] = int(near.sum())
    out["hard"] = int((~near).sum())
    out["max_gap"] = float(gap.max())
    return out
