"""mu-law decode table used by the sample loop's epilogue (formula of /root/reference/preprocess.py:30-35)."""
from __future__ import annotations

import numpy as np


def mulaw_decode_table(bits: int = 8) -> np.ndarray:
    """k in [0, 2**bits) -> linear amplitude: y = 2k/(mu-1) - 1 ; x = sign(y)/(mu-1) * (mu**|y| - 1)."""
    mu = 1 << bits
    y = 2.0 * np.arange(mu, dtype=np.float64) / (mu - 1) - 1.0
    return (np.sign(y) / (mu - 1) * (float(mu) ** np.abs(y) - 1.0)).astype(np.float32)
