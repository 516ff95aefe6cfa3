"""ORACLE (test infrastructure -- never imported by the product): CPU restatement of the reference's log-mel
front-end, ``wave_to_mel`` at /root/reference/preprocess.py:53-75 (same arithmetic inline in convert.py:54-70).

The reference delegates the STFT, the mel filterbank and the dB conversion to ``librosa ^0.8.0`` (pyproject.toml:18),
a third-party dependency that is ABSENT from this image; its published algorithm is restated here in float64 numpy:

  librosa.feature.melspectrogram(y, sr, n_fft, hop_length, win_length, window='hann', center=True, pad_mode='reflect',
                                 power=1, n_mels, fmin)    = mel_basis . |stft(y)|
  librosa.stft: reflect-pad n_fft//2 on both sides, frames of n_fft at multiples of hop_length, periodic hann(win_length)
                zero-padded to n_fft (centred), rfft
  librosa.filters.mel(htk=False, norm='slaney'): triangles on the Slaney mel scale, each scaled by 2 / (f[i+2] - f[i])
  librosa.amplitude_to_db(S, ref=1.0, amin=1e-5, top_db): 20 log10(max(amin, S)) clamped to (max - top_db)

PARITY STATUS: pinned to torchaudio's librosa-compatible ``MelSpectrogram`` (norm='slaney', mel_scale='slaney',
pad_mode='reflect', power=1) and to scipy.signal.stft for the framing (tests/test_frontend_cpu.py); NOT pinned to
librosa itself (absent here) -- say "parity unpinned against librosa" wherever this front-end is quoted."""
import math

import numpy as np


def preemphasis(x: np.ndarray, preemph: float) -> np.ndarray:
    """preprocess.py:16-17: scipy.signal.lfilter([1, -preemph], [1], x) with a zero initial state."""
    x = np.asarray(x, dtype=np.float64)
    y = x.copy()
    y[1:] -= preemph * x[:-1]
    return y


def hann_periodic(n: int) -> np.ndarray:
    return np.array([0.5 - 0.5 * math.cos(2.0 * math.pi * i / n) for i in range(n)])


def stft_magnitude(y: np.ndarray, n_fft: int, hop: int, win: int) -> np.ndarray:
    """|librosa.stft(y, n_fft, hop, win, 'hann', center=True, pad_mode='reflect')| -> (1 + n_fft//2, 1 + len(y)//hop)."""
    y = np.asarray(y, dtype=np.float64)
    pad = n_fft // 2
    yp = np.pad(y, pad, mode="reflect")
    w = np.zeros(n_fft)
    lo = (n_fft - win) // 2
    w[lo:lo + win] = hann_periodic(win)
    n_frames = 1 + (len(yp) - n_fft) // hop
    out = np.empty((n_fft // 2 + 1, n_frames))
    for t in range(n_frames):
        out[:, t] = np.abs(np.fft.rfft(w * yp[t * hop:t * hop + n_fft]))
    return out


def _hz_to_mel(f: float) -> float:
    f_sp = 200.0 / 3
    if f >= 1000.0:
        return 1000.0 / f_sp + math.log(f / 1000.0) / (math.log(6.4) / 27.0)
    return f / f_sp


def _mel_to_hz(m: float) -> float:
    f_sp = 200.0 / 3
    if m >= 1000.0 / f_sp:
        return 1000.0 * math.exp((math.log(6.4) / 27.0) * (m - 1000.0 / f_sp))
    return f_sp * m


def mel_basis(sr: int, n_fft: int, n_mels: int, fmin: float) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax=sr/2, htk=False, norm='slaney') -> (n_mels, 1 + n_fft//2)."""
    fmax = sr / 2.0
    m_lo, m_hi = _hz_to_mel(fmin), _hz_to_mel(fmax)
    edges = [_mel_to_hz(m_lo + (m_hi - m_lo) * i / (n_mels + 1)) for i in range(n_mels + 2)]
    n_freq = n_fft // 2 + 1
    freqs = [sr / 2.0 * k / (n_freq - 1) for k in range(n_freq)]
    W = np.zeros((n_mels, n_freq))
    for i in range(n_mels):
        lo, ce, hi = edges[i], edges[i + 1], edges[i + 2]
        for k, f in enumerate(freqs):
            up, down = (f - lo) / (ce - lo), (hi - f) / (hi - ce)
            W[i, k] = max(0.0, min(up, down)) * 2.0 / (hi - lo)
    return W


def wave_to_mel(wave: np.ndarray, sr=16000, n_fft=2048, n_mels=80, fmin=50, preemph=0.97, top_db=80, hop_length=160,
                win_length=400) -> np.ndarray:
    """preprocess.py:53-75 -> (n_mels, 1 + len(wave)//hop) float64."""
    wave = np.asarray(wave, dtype=np.float64)
    wave_s = wave / np.abs(wave).max() * 0.999                                  # preprocess.py:62
    S = stft_magnitude(preemphasis(wave_s, preemph), n_fft, hop_length, win_length)
    mel = mel_basis(sr, n_fft, n_mels, fmin) @ S                                # power = 1
    db = 20.0 * np.log10(np.maximum(1e-5, mel))                                 # amplitude_to_db, ref = 1
    db = np.maximum(db, db.max() - top_db)
    return db / top_db + 1.0                                                    # preprocess.py:74
