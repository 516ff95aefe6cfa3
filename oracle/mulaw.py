"""mu-law companding -- TEST INFRASTRUCTURE (see oracle/__init__.py).  PINNED against
/root/reference/preprocess.py:20-35 (``tests/golden/mulaw_lut.npz``)."""
from __future__ import annotations

import numpy as np


def mulaw_decode_float(y: np.ndarray, mu: int = 256) -> np.ndarray:
    """/root/reference/preprocess.py:30-35: x = sign(y)/(mu-1) * ((1+(mu-1))**|y| - 1), y in [-1,1]."""
    m = mu - 1
    y = np.asarray(y, dtype=np.float64)
    return np.sign(y) / m * ((1.0 + m) ** np.abs(y) - 1.0)


def mulaw_encode_float(x: np.ndarray, mu: int = 256) -> np.ndarray:
    """/root/reference/preprocess.py:20-27: floor((sign(x) log1p(m|x|)/log1p(m) + 1)/2 * m + 0.5), m = mu-1."""
    m = mu - 1
    x = np.asarray(x, dtype=np.float64)
    fx = np.sign(x) * np.log1p(m * np.abs(x)) / np.log1p(m)
    return np.floor((fx + 1) / 2 * m + 0.5)


def mulaw_decode_lut(bits: int = 8) -> np.ndarray:
    """Table k -> wav for k in [0, 2**bits): y = 2k/(mu-1) - 1 then ``mulaw_decode_float``.
    Evaluated in float64, returned as float32 (256 entries; code 128 -> +8.62e-5, 127 -> -8.62e-5)."""
    mu = 1 << bits
    k = np.arange(mu, dtype=np.float64)
    y = 2.0 * k / (mu - 1) - 1.0
    return mulaw_decode_float(y, mu).astype(np.float32)
