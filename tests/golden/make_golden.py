#!/usr/bin/env python
"""Generate tests/golden/*.npz from the LIVE reference (run in the build container only).

    python tests/golden/make_golden.py

Reads /root/reference (read-only) through oracle/reference_loader.py: the genuine ``Encoder`` /
``VQEmbeddingEMA`` classes from model.py:1-317 and the two pure-numpy mu-law functions from
preprocess.py:20-35.  Writes OUTPUTS only (plus seeds and a weight checksum): weights and inputs are
re-created from seeds by ``oracle/fixtures.py`` at test time, so the fixtures stay small.

The vocoder core (``rnnms``) is not in the reference tree -> no reference-generated fixture exists for
it ("parity unpinned", see oracle/__init__.py).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import fixtures, reference_loader  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

ENCODER_CASES = [
    # name, channels, B, T, mel kind, mel seed, perturbed weights?
    ("encoder_c768_T200_init", 768, 1, 200, "rand", 0, False),      # BASELINE config 1
    ("encoder_c512_T201_trained", 512, 1, 201, "rand", 1, True),    # reference default width, odd T
    ("encoder_c768_T301_B3_trained", 768, 3, 301, "randn", 2, True),
    ("encoder_c768_T8_B2_trained", 768, 2, 8, "rand", 3, True),     # tiny / edge
]


def make_encoder():
    for name, C, B, T, kind, mseed, perturbed in ENCODER_CASES:
        enc = reference_loader.build_reference_encoder(C, seed=13)
        sd = {k: v.clone() for k, v in enc.state_dict().items()}
        if perturbed:
            sd = fixtures.perturb_encoder_state(sd)
            enc.load_state_dict(sd)
        mel = fixtures.synthetic_mel(B, T, seed=mseed, kind=kind)
        aux = []
        handle = enc.encoder[-1].register_forward_hook(lambda m, i, o: aux.append(o.clone()))
        with torch.no_grad():
            z, c, idx = enc.encode(mel)
        handle.remove()
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"),
            channels=C, B=B, T=T, mel_kind=kind, mel_seed=mseed, perturbed=perturbed, weight_seed=13,
            weight_checksum=fixtures.state_checksum(sd),
            z=z.numpy(), c=c.numpy(), indices=idx.numpy(), z_pre=aux[0].numpy(),
        )
        print(name, tuple(z.shape), tuple(c.shape), tuple(idx.shape))


def make_vq():
    m = reference_loader.load_reference_model()
    out = {}
    for kind in ("init", "trained"):
        x, cb = fixtures.vq_inputs(20000, kind=kind, seed=1234, batch=2)
        vq = m.VQEmbeddingEMA(512, 64)
        vq.embedding.copy_(cb)
        with torch.no_grad():
            q, idx = vq.encode(x)
        out[f"{kind}_indices"] = idx.numpy().astype(np.int16)
        out[f"{kind}_qsum"] = float(q.double().sum())
    # exact tie: duplicate codes 7 and 300 -> lowest index wins (torch.argmin)
    x, cb = fixtures.vq_inputs(16, kind="trained", seed=5)
    cb[300] = cb[7]
    xq = cb[[7, 300, 12, 7]][None]
    vq = m.VQEmbeddingEMA(512, 64)
    vq.embedding.copy_(cb)
    with torch.no_grad():
        _, idx = vq.encode(xq)
    out["tie_indices"] = idx.numpy()
    np.savez_compressed(os.path.join(OUT, "vq_lookup.npz"), **out)
    print("vq_lookup", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


def make_mulaw():
    enc, dec = reference_loader.load_reference_preprocess_mulaw()
    k = np.arange(256, dtype=np.float64)
    lut = dec(2.0 * k / 255.0 - 1.0, 256)
    grid = np.linspace(-1.0, 1.0, 2001)
    codes = enc(grid, 256)
    np.savez_compressed(os.path.join(OUT, "mulaw_lut.npz"), decode_lut=lut, encode_grid=grid, encode_codes=codes)
    print("mulaw_lut", lut[[0, 127, 128, 255]])


if __name__ == "__main__":
    make_encoder()
    make_vq()
    make_mulaw()
