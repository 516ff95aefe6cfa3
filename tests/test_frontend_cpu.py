"""Pins the front-end oracle (oracle/frontend.py) -- librosa is absent here, so the pins are torchaudio's librosa-
compatible MelSpectrogram, scipy's STFT and scipy's lfilter -- and checks the product's host-side constant matrices."""
import numpy as np
import scipy.signal
import torch
import torchaudio

from oracle import frontend as ofe
from vectorquantizedcpc_b200 import frontend as fe


def _signal(n=8000, seed=0):
    g = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    return (0.3 * np.sin(2 * np.pi * 220 * t) + 0.2 * np.sin(2 * np.pi * 3100 * t * (1 + 0.2 * t)) + 0.05 * g.standard_normal(n)).astype(np.float32)


def test_preemphasis_is_scipy_lfilter():
    x = _signal(4000, 1).astype(np.float64)
    assert np.allclose(ofe.preemphasis(x, 0.97), scipy.signal.lfilter([1, -0.97], [1], x), rtol=0, atol=1e-15)   # preprocess.py:16-17


def test_stft_framing_matches_scipy():
    y = _signal(8000, 2).astype(np.float64)
    S = ofe.stft_magnitude(y, 2048, 160, 400)
    assert S.shape == (1025, 1 + 8000 // 160)
    # scipy centres the 400-tap window on t*hop with an even (reflect) extension: the same 400 samples per frame
    _, _, Z = scipy.signal.stft(y, window=scipy.signal.get_window("hann", 400), nperseg=400, noverlap=240, nfft=2048,
                                boundary="even", padded=False, scaling="spectrum")
    ref = np.abs(Z) * scipy.signal.get_window("hann", 400).sum()
    assert ref.shape == S.shape
    assert np.allclose(S, ref, rtol=1e-9, atol=1e-9)


def test_mel_basis_matches_torchaudio_slaney():
    ref = torchaudio.functional.melscale_fbanks(n_freqs=1025, f_min=50.0, f_max=8000.0, n_mels=80, sample_rate=16000,
                                                norm="slaney", mel_scale="slaney").T.numpy()
    W = ofe.mel_basis(16000, 2048, 80, 50)
    assert W.shape == (80, 1025)
    assert np.allclose(W, ref, rtol=2e-4, atol=1e-7)
    P = fe.mel_filterbank(16000, 2048, 80, 50)              # the product's own construction
    assert np.allclose(P, W, rtol=1e-10, atol=1e-14)


def test_melspectrogram_matches_torchaudio():
    x = _signal(16000, 3)
    y = ofe.preemphasis(x.astype(np.float64) / np.abs(x).max() * 0.999, 0.97)
    ms = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=2048, win_length=400, hop_length=160, f_min=50.0,
                                              n_mels=80, power=1.0, center=True, pad_mode="reflect", norm="slaney",
                                              mel_scale="slaney")
    ref = ms(torch.from_numpy(y).float()[None])[0].double().numpy()
    mel = ofe.mel_basis(16000, 2048, 80, 50) @ ofe.stft_magnitude(y, 2048, 160, 400)
    assert mel.shape == ref.shape == (80, 101)
    assert np.allclose(mel, ref, rtol=2e-3, atol=2e-4 * ref.max())      # torchaudio runs in fp32


def test_wave_to_mel_range_and_clamp():
    m = ofe.wave_to_mel(_signal(16000, 4))
    assert m.shape == (80, 101)
    assert abs(m.max() - (20 * np.log10(np.maximum(1e-5, (ofe.mel_basis(16000, 2048, 80, 50) @ ofe.stft_magnitude(
        ofe.preemphasis(_signal(16000, 4).astype(np.float64) / np.abs(_signal(16000, 4)).max() * 0.999, 0.97), 2048, 160, 400)))).max() / 80 + 1)) < 1e-12
    assert m.min() >= m.max() - 1.0 - 1e-12                                # top_db / top_db = 1 below the maximum


def test_product_constant_matrices():
    lm = fe.LogMel()
    assert lm.dft.shape == (2 * 1040, 400) and lm.melw.shape == (80, 1040) and lm.window.shape == (400,)
    assert np.allclose(lm.window.numpy(), scipy.signal.get_window("hann", 400), atol=1e-7)
    assert not lm.dft[1025:1040].any() and not lm.dft[1040 + 1025:].any() and not lm.melw[:, 1025:].any()
    # row k of the cos / sin blocks against a direct DFT of a windowed frame
    g = np.random.default_rng(0)
    a = g.standard_normal(400)
    frame = np.zeros(2048); frame[824:1224] = a
    spec = np.abs(np.fft.rfft(frame))
    re, im = lm.dft[:1025].double().numpy() @ a, lm.dft[1040:1040 + 1025].double().numpy() @ a
    assert np.allclose(np.hypot(re, im), spec, rtol=1e-4, atol=1e-4)


def test_loudness_oracle_matches_torchaudio_bs1770():
    """oracle/loudness.py restates pyloudnorm (absent here); torchaudio.functional.loudness is an independent BS.1770-4
    implementation -- agreement within 0.05 LU on signals long enough that one trailing block does not matter."""
    from oracle import loudness as olo
    for seed, scale in ((5, 1.0), (6, 0.1), (7, 0.01)):
        x = _signal(48000, seed) * scale
        ref = float(torchaudio.functional.loudness(torch.from_numpy(x)[None], 16000))
        got = olo.integrated_loudness(x, 16000)
        assert abs(got - ref) < 0.05, (got, ref)
    a = olo.integrated_loudness(_signal(48000, 5), 16000)
    assert abs(olo.integrated_loudness(_signal(48000, 5) * 0.5, 16000) - (a + 20 * np.log10(0.5))) < 1e-9     # linearity
    y = olo.normalize_loudness(_signal(48000, 5), a, -23.0)
    assert abs(olo.integrated_loudness(y, 16000) + 23.0) < 1e-9
    assert olo.integrated_loudness(np.zeros(16000), 16000) == -np.inf
