"""Output stage of ``convert.py`` on the GPU: pyloudnorm's integrated loudness
(``pyloudnorm.Meter(sr).integrated_loudness``, convert.py:50,57,79) and gain (``pyloudnorm.normalize.loudness``,
convert.py:80), so a converted batch can stay on the device until it is written out.  Kernels: ``csrc/loudness.cu``."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib


def _prep(wave: Tensor, lengths) -> Tuple[Tensor, Optional[Tensor]]:
    _lib.require_cuda(wave, "wave")
    if wave.dim() == 1:
        wave = wave[None]
    if wave.dim() != 2 or wave.dtype != torch.float32:
        raise ValueError("wave must be (B, N) or (N,) float32 (mono)")
    len_t = None
    if lengths is not None:
        len_t = torch.as_tensor(lengths)
        if len_t.dim() != 1 or len_t.shape[0] != wave.shape[0] or len_t.dtype.is_floating_point:
            raise ValueError("lengths must be an integer vector with one entry per utterance")
        if len_t.numel() and (int(len_t.min()) < 0 or int(len_t.max()) > wave.shape[1]):
            raise ValueError(f"lengths must lie in [0, {wave.shape[1]}]")
        len_t = len_t.to(device=wave.device, dtype=torch.int32).contiguous()
    return wave.contiguous(), len_t


def integrated_loudness(wave: Tensor, rate: int = 16000, lengths=None) -> Tensor:
    """``pyloudnorm.Meter(rate).integrated_loudness(wave[b])`` for every utterance -> (B,) fp32 LUFS (-inf for silence).
    The signal must be longer than one 400 ms block, as pyloudnorm requires."""
    wave, len_t = _prep(wave, lengths)
    B, N = wave.shape
    shortest = N if len_t is None or B == 0 else int(len_t.min())
    if B and shortest <= int(0.4 * rate):
        raise ValueError("Audio must have length greater than the block size (0.4 s)")       # pyloudnorm's own check
    out = torch.empty(B, device=wave.device)
    if B == 0:
        return out
    lib = _lib.lib()
    ws_bytes = lib.vqcpc_loudness_workspace_bytes(B, N, rate)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=wave.device)
    with torch.cuda.device(wave.device):
        _lib.check(lib.vqcpc_integrated_loudness(_lib.ptr(wave), _lib.ptr(len_t), B, N, rate, _lib.ptr(ws), ws_bytes, _lib.ptr(out),
                                                 _lib.current_stream_ptr()), "integrated_loudness")
    return out


def loudness_normalize(wave: Tensor, target_lufs, rate: int = 16000, lengths=None) -> Tuple[Tensor, Tensor]:
    """convert.py:79-80 in one call: measure every utterance, then scale it to ``target_lufs[b]`` (the loudness of the
    source utterance, convert.py:57).  Returns (normalised (B, N), measured LUFS (B,))."""
    wave, len_t = _prep(wave, lengths)
    B, N = wave.shape
    target = torch.as_tensor(target_lufs, dtype=torch.float32, device=wave.device).reshape(-1).contiguous()
    if target.shape[0] != B:
        raise ValueError("target_lufs must have one entry per utterance")
    shortest = N if len_t is None or B == 0 else int(len_t.min())
    if B and shortest <= int(0.4 * rate):
        raise ValueError("Audio must have length greater than the block size (0.4 s)")
    out, measured = torch.empty_like(wave), torch.empty(B, device=wave.device)
    if B == 0:
        return out, measured
    lib = _lib.lib()
    ws_bytes = lib.vqcpc_loudness_workspace_bytes(B, N, rate)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=wave.device)
    with torch.cuda.device(wave.device):
        _lib.check(lib.vqcpc_loudness_normalize(_lib.ptr(wave), _lib.ptr(len_t), _lib.ptr(target), B, N, rate, _lib.ptr(ws), ws_bytes,
                                                _lib.ptr(out), _lib.ptr(measured), _lib.current_stream_ptr()), "loudness_normalize")
    return out, measured
