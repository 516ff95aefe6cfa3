"""Log-mel front-end on the GPU -- the step immediately before ``Encoder.encode``.

Mirror of ``wave_to_mel`` (/root/reference/preprocess.py:53-75; the same arithmetic is inline in convert.py:54-70):
peak scaling, pre-emphasis, magnitude STFT, Slaney mel projection, dB with a per-utterance ``top_db`` clamp, ``/top_db + 1``.
The host builds three small constant matrices once (window, [cos | sin] DFT rows of the window's non-zero taps, mel
filterbank); every per-sample operation runs in ``csrc/frontend.cu`` behind ``vqcpc_logmel_forward``.

The reference computes this with librosa (``^0.8.0``, pyproject.toml:18) which is not in this image, so parity is
pinned against torchaudio's librosa-compatible ``MelSpectrogram`` and scipy's STFT instead (tests/test_frontend_cpu.py,
oracle/frontend.py)."""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
from torch import Tensor, nn

from . import _lib


@dataclass
class ConfPreprocessing:
    """preprocess.py:39-51; defaults = config.py:103-112 (sampling_rate 16000, dim_mel_freq 80, hop 160)."""
    sr: int = 16000
    n_fft: int = 2048
    n_mels: int = 80
    fmin: int = 50
    preemph: float = 0.97
    top_db: int = 80
    hop_length: int = 160
    win_length: int = 400
    bits: int = 8


def _hz_to_mel_slaney(f: np.ndarray) -> np.ndarray:
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, logstep = 1000.0, math.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz_slaney(m: np.ndarray) -> np.ndarray:
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, logstep = 1000.0, math.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: Optional[float] = None) -> np.ndarray:
    """(n_mels, n_fft//2 + 1) triangular filters on the Slaney mel scale with Slaney area normalisation -- what
    ``librosa.filters.mel(sr, n_fft, n_mels, fmin)`` (htk=False, norm='slaney') builds for melspectrogram."""
    fmax = sr / 2.0 if fmax is None else float(fmax)
    fft_f = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    mel_f = _mel_to_hz_slaney(np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_f[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w


class LogMel(nn.Module):
    """``LogMel(conf)(wave)``: wave ``(B, N)`` or ``(N,)`` fp32 on the GPU -> log-mel ``(B, n_mels, 1 + N // hop)`` fp32,
    the tensor ``Encoder.encode`` takes.  ``lengths`` (B,) = valid samples per utterance of a zero-padded batch: every
    utterance is scaled, framed, reflected and clamped on its own, frames beyond ``1 + lengths[b] // hop`` are 0."""

    def __init__(self, conf: Optional[ConfPreprocessing] = None, **kwargs):
        super().__init__()
        if conf is None:
            conf = ConfPreprocessing(**kwargs)
        elif kwargs:
            raise TypeError("pass either a ConfPreprocessing or keyword fields, not both")
        if conf.win_length % 16 or conf.win_length > conf.n_fft or conf.n_fft % 2 or conf.n_mels % 4:
            raise ValueError("unsupported: win_length must be a multiple of 16 and <= n_fft (even), n_mels a multiple of 4")
        self.conf = conf
        n_freq = conf.n_fft // 2 + 1
        nfp = (n_freq + 15) // 16 * 16
        self.n_freq_padded = nfp
        j = np.arange(conf.win_length, dtype=np.float64)
        window = 0.5 - 0.5 * np.cos(2.0 * np.pi * j / conf.win_length)          # periodic hann (scipy get_window, fftbins=True)
        # the window sits centred in the n_fft frame; its offset only rotates the phase, which |.| drops
        ang = 2.0 * np.pi * np.outer(np.arange(n_freq, dtype=np.float64), j) / conf.n_fft
        dft = np.zeros((2 * nfp, conf.win_length), dtype=np.float64)
        dft[:n_freq] = np.cos(ang)
        dft[nfp:nfp + n_freq] = np.sin(ang)
        melw = np.zeros((conf.n_mels, nfp), dtype=np.float64)
        melw[:, :n_freq] = mel_filterbank(conf.sr, conf.n_fft, conf.n_mels, conf.fmin)
        self.register_buffer("window", torch.from_numpy(window.astype(np.float32)), persistent=False)
        self.register_buffer("dft", torch.from_numpy(dft.astype(np.float32)), persistent=False)
        self.register_buffer("melw", torch.from_numpy(melw.astype(np.float32)), persistent=False)

    def _cfg(self) -> _lib.LogMelConfig:
        c = self.conf
        return _lib.LogMelConfig(c.n_fft, c.win_length, c.hop_length, c.n_mels, self.n_freq_padded, float(c.preemph), float(c.top_db))

    def forward(self, wave: Tensor, lengths=None) -> Tensor:
        _lib.require_cuda(wave, "wave")
        if wave.dim() == 1:
            wave = wave[None]
        if wave.dim() != 2 or wave.dtype != torch.float32:
            raise ValueError("wave must be (B, N) or (N,) float32")
        if self.window.device != wave.device:
            raise RuntimeError("LogMel buffers and wave are on different devices: call .to(wave.device) first")
        B, N = wave.shape
        half = self.conf.n_fft // 2
        if B and N <= half:
            raise ValueError(f"need more than n_fft/2 = {half} samples per utterance (reflect padding), got {N}")
        len_t = None
        if lengths is not None:
            len_t = torch.as_tensor(lengths)
            if len_t.dim() != 1 or len_t.shape[0] != B or len_t.dtype.is_floating_point:
                raise ValueError("lengths must be an integer vector with one entry per utterance")
            if B and (int(len_t.min()) <= half or int(len_t.max()) > N):
                raise ValueError(f"lengths must lie in ({half}, {N}]")
            len_t = len_t.to(device=wave.device, dtype=torch.int32).contiguous()
        T = 1 + N // self.conf.hop_length
        out = torch.empty(B, self.conf.n_mels, T, device=wave.device)
        if B == 0:
            return out
        lib = _lib.lib()
        cfg = self._cfg()
        ws_bytes = lib.vqcpc_logmel_workspace_bytes(C.byref(cfg), B, N)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=wave.device)
        wc = wave.contiguous()
        with torch.cuda.device(wave.device):
            _lib.check(lib.vqcpc_logmel_forward(C.byref(cfg), _lib.ptr(wc), _lib.ptr(len_t), B, N, _lib.ptr(self.window),
                                                _lib.ptr(self.dft), _lib.ptr(self.melw), _lib.ptr(ws), ws_bytes, _lib.ptr(out),
                                                _lib.current_stream_ptr()), "LogMel")
        return out


def wave_to_mel(wave: Tensor, conf: Optional[ConfPreprocessing] = None) -> Tensor:
    """Functional form with the reference's name (preprocess.py:53): one utterance ``(N,)`` -> ``(n_mels, T)``."""
    fe = LogMel(conf).to(wave.device)
    return fe(wave)[0]
