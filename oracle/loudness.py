"""ORACLE (test infrastructure -- never imported by the product): CPU restatement of the loudness calls of
/root/reference/convert.py:50,57,79-80 -- ``pyloudnorm.Meter(sr).integrated_loudness`` and
``pyloudnorm.normalize.loudness``.

pyloudnorm is a third-party dependency ABSENT from this image (unpinned in the reference: imported by convert.py:9,
not listed in pyproject.toml / requirements.txt); its published algorithm (ITU-R BS.1770-4 with the "K-weighting"
filter class: RBJ high shelf +4 dB / Q 1/sqrt2 / 1500 Hz and high pass Q 0.5 / 38 Hz derived for the given rate,
400 ms blocks with 75 % overlap, -70 LUFS absolute gate, -10 LU relative gate) is restated here.

PARITY STATUS: pinned to torchaudio.functional.loudness (an independent BS.1770-4 implementation) within 0.05 LU
(tests/test_frontend_cpu.py); NOT pinned to pyloudnorm itself -- "parity unpinned against pyloudnorm"."""
import numpy as np
import scipy.signal


def _coefficients(G, Q, fc, rate, kind):
    A = 10 ** (G / 40.0)
    w0 = 2.0 * np.pi * (fc / rate)
    alpha = np.sin(w0) / (2.0 * Q)
    if kind == "high_shelf":
        b0 = A * ((A + 1) + (A - 1) * np.cos(w0) + 2 * np.sqrt(A) * alpha)
        b1 = -2 * A * ((A - 1) + (A + 1) * np.cos(w0))
        b2 = A * ((A + 1) + (A - 1) * np.cos(w0) - 2 * np.sqrt(A) * alpha)
        a0 = (A + 1) - (A - 1) * np.cos(w0) + 2 * np.sqrt(A) * alpha
        a1 = 2 * ((A - 1) - (A + 1) * np.cos(w0))
        a2 = (A + 1) - (A - 1) * np.cos(w0) - 2 * np.sqrt(A) * alpha
    else:
        b0 = (1 + np.cos(w0)) / 2
        b1 = -(1 + np.cos(w0))
        b2 = (1 + np.cos(w0)) / 2
        a0 = 1 + alpha
        a1 = -2 * np.cos(w0)
        a2 = 1 - alpha
    return np.array([b0, b1, b2]) / a0, np.array([a0, a1, a2]) / a0


def integrated_loudness(data: np.ndarray, rate: int) -> float:
    """Mono ``Meter(rate).integrated_loudness(data)``."""
    x = np.asarray(data, dtype=np.float64)
    for G, Q, fc, kind in ((4.0, 1 / np.sqrt(2), 1500.0, "high_shelf"), (0.0, 0.5, 38.0, "high_pass")):
        b, a = _coefficients(G, Q, fc, rate, kind)
        x = scipy.signal.lfilter(b, a, x)
    T_g, Gamma_a, step = 0.4, -70.0, 0.25
    T = x.shape[0] / rate
    n_blocks = int(np.round(((T - T_g) / (T_g * step))) + 1)
    z = np.zeros(n_blocks)
    for j in range(n_blocks):
        lo = int(T_g * (j * step) * rate)
        hi = int(T_g * (j * step + 1) * rate)
        z[j] = (1.0 / (T_g * rate)) * np.sum(np.square(x[lo:hi]))
    with np.errstate(divide="ignore", invalid="ignore"):
        l = -0.691 + 10.0 * np.log10(z)
        J = [j for j in range(n_blocks) if l[j] >= Gamma_a]
        Gamma_r = -0.691 + 10.0 * np.log10(np.mean(z[J])) - 10.0 if J else np.nan
        J = [j for j in range(n_blocks) if l[j] > Gamma_r and l[j] > Gamma_a]
        z_avg = np.nan_to_num(np.mean(z[J])) if J else 0.0
        return float(-0.691 + 10.0 * np.log10(z_avg))


def normalize_loudness(data: np.ndarray, input_loudness: float, target_loudness: float) -> np.ndarray:
    """``pyloudnorm.normalize.loudness``: gain = 10^((target - input) / 20)."""
    return np.asarray(data, dtype=np.float64) * np.power(10.0, (target_loudness - input_loudness) / 20.0)
