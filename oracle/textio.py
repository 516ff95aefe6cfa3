"""TEST INFRASTRUCTURE -- the reference's own file writers, and a plain-Python restatement of the "%.16f" arithmetic.

* ``savetxt_bytes``  = ``np.savetxt(f, a, fmt="%.16f")`` (/root/reference/encode.py:51,59,66) -- numpy itself, so the GPU text
  dump is pinned to the function the reference calls, not to a restatement.
* ``wav_bytes_scipy`` = ``scipy.io.wavfile.write`` of a float32 array, which is what ``librosa.output.write_wav`` (librosa
  0.7.x, /root/reference/convert.py:83) runs after its validity checks.
* ``format_f16`` restates csrc/textdump.cu's integer arithmetic with Python integers; the CPU suite checks it against
  ``"%.16f" % v`` on edge cases and random bit patterns, which validates the algorithm where no GPU is present."""
from __future__ import annotations

import io
import struct

import numpy as np


def savetxt_bytes(a) -> bytes:
    buf = io.BytesIO()
    np.savetxt(buf, np.asarray(a), fmt="%.16f")
    return buf.getvalue()


def wav_bytes_scipy(y, sr: int) -> bytes:
    from scipy.io import wavfile
    buf = io.BytesIO()
    wavfile.write(buf, sr, np.asarray(y, dtype=np.float32))
    return buf.getvalue()


def format_f16(v) -> str:
    """"%.16f" % float(np.float32(v)) by the integer arithmetic of csrc/textdump.cu."""
    bits = struct.unpack("<I", struct.pack("<f", float(np.float32(v))))[0] if not isinstance(v, (int, np.integer)) else int(v)
    neg = bits >> 31
    ex, man = (bits >> 23) & 0xFF, bits & 0x7FFFFF
    if ex == 0xFF:
        return "nan" if man else ("-inf" if neg else "inf")
    m = (man | 0x800000) if ex else man
    e = ex - 150 if ex else -149
    if e >= 0:
        ipart, D = m << e, 0
    else:
        q = -e
        ipart, k = (m >> q, m & ((1 << q) - 1)) if q < 32 else (0, m)
        if q <= 16:
            D = (k * 5 ** 16) << (16 - q)
        else:
            N, sh = k * 5 ** 16, q - 16
            if sh > 62:
                D = 0
            else:
                D, rem, half = N >> sh, N & ((1 << sh) - 1), 1 << (sh - 1)
                if rem > half or (rem == half and (D & 1)):
                    D += 1
        if D >= 10 ** 16:
            D -= 10 ** 16
            ipart += 1
    return ("-" if neg else "") + str(ipart) + "." + str(D).rjust(16, "0")
