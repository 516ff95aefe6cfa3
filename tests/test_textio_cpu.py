"""CPU checks of the output-file writers: the "%.16f" integer arithmetic of csrc/textdump.cu restated in
oracle/textio.py against Python's own formatting, and the wav container against scipy.io.wavfile.write (what
librosa.output.write_wav of convert.py:83 calls)."""
import io

import numpy as np
import torch

from oracle import textio as otext
from vectorquantizedcpc_b200 import textio


def test_fixed16_integer_arithmetic_equals_printf():
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2 ** 32, size=30000, dtype=np.uint64).astype(np.uint32)
    # plus values of the sizes the encoder writes (|z| < 10) and exact ties at the 17th digit (k odd / 2^17)
    small = (rng.standard_normal(5000) * 2).astype(np.float32).view(np.uint32)
    ties = (np.arange(1, 400, 2, dtype=np.float32) * np.float32(2.0 ** -17)).view(np.uint32)
    edge = np.array([0.0, -0.0, 1.0, -1.0, 0.5, 0.99999994, 9.9999999, 2.0 ** -149, 1e-45, 2.0 ** 24, 3.4e38, -3.4e38, np.inf, -np.inf,
                     2.0 ** -16, 2.0 ** -17, 3 * 2.0 ** -18], dtype=np.float32).view(np.uint32)
    for b in np.concatenate([bits, small, ties, edge]):
        v = np.array([b], dtype=np.uint32).view(np.float32)[0]
        if np.isnan(v):
            assert otext.format_f16(int(b)) == "nan"
            continue
        assert otext.format_f16(int(b)) == "%.16f" % float(v), hex(int(b))


def test_savetxt_oracle_layout():
    a = np.array([[1.5, -2.25, 0.0], [3.0, 4.0, -0.0]], dtype=np.float32)
    want = b"1.5000000000000000 -2.2500000000000000 0.0000000000000000\n3.0000000000000000 4.0000000000000000 -0.0000000000000000\n"
    assert otext.savetxt_bytes(a) == want
    assert " ".join(otext.format_f16(v) for v in a[0]).encode() + b"\n" == want.split(b"\n")[0] + b"\n"


def test_wav_container_equals_scipy():
    rng = np.random.default_rng(0)
    for shape in [(16000,), (1,), (0,), (4001, 2)]:
        y = rng.standard_normal(shape).astype(np.float32) * 0.3
        assert textio.wav_bytes(torch.from_numpy(y), 16000) == otext.wav_bytes_scipy(y, 16000), shape
    # float64 input is written as float32, as convert.py:83 does with .astype(np.float32)
    y64 = rng.standard_normal(100)
    assert textio.wav_bytes(torch.from_numpy(y64), 22050) == otext.wav_bytes_scipy(y64.astype(np.float32), 22050)


def test_write_wav_roundtrip(tmp_path):
    from scipy.io import wavfile
    y = (np.sin(np.arange(8000) * 0.05) * 0.5).astype(np.float32)
    p = tmp_path / "a.wav"
    textio.write_wav(p, torch.from_numpy(y), sr=16000)
    sr, back = wavfile.read(p)
    assert sr == 16000 and back.dtype == np.float32 and np.array_equal(back, y)
