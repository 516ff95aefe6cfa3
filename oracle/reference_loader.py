"""Load the LIVE reference ``Encoder`` / ``VQEmbeddingEMA`` (build container only).

TEST INFRASTRUCTURE.  ``/root/reference`` exists only in the build container,
never on the GPU box: this loader is used by ``tests/golden/make_golden.py``
(fixture generation) and by CPU tests that are skipped when the tree is absent.

The reference does not import as checked in:
  * ``model.py:6`` needs ``omegaconf.omegaconf.MISSING`` (omegaconf is not installed);
  * ``model.py:319-322`` (``ConfModel``) raises ``ValueError: mutable default`` on
    Python >= 3.11.
So we exec source lines 1..317 only -- everything above ``ConfModel`` -- into a
fresh module, with a one-attribute ``omegaconf`` stub.  No reference source is
copied into this repository; it is read from where it lies.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VQCPC_REFERENCE_ROOT", "/root/reference")
_N_LINES = 317  # model.py:1-317 = Encoder, VQEmbeddingEMA, ConfCPC, CPCLoss


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model.py"))


def load_reference_model() -> types.ModuleType:
    """Return a module object holding the reference's ``Encoder``, ``VQEmbeddingEMA``, ``ConfEncoder``."""
    if "_vqcpc_reference_model" in sys.modules:
        return sys.modules["_vqcpc_reference_model"]
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    if "omegaconf" not in sys.modules:
        stub = types.ModuleType("omegaconf")
        stub.MISSING = "???"
        sub = types.ModuleType("omegaconf.omegaconf")
        sub.MISSING = "???"
        stub.omegaconf = sub
        sys.modules["omegaconf"] = stub
        sys.modules["omegaconf.omegaconf"] = sub
    path = os.path.join(REFERENCE_ROOT, "model.py")
    with open(path, "r") as f:
        lines = f.readlines()[:_N_LINES]
    mod = types.ModuleType("_vqcpc_reference_model")
    mod.__file__ = path
    sys.modules["_vqcpc_reference_model"] = mod
    exec(compile("".join(lines), path, "exec"), mod.__dict__)
    return mod


def build_reference_encoder(channels: int = 768, seed: int = 13):
    """Reference ``Encoder`` with default ``nn.Module`` init under ``seed`` (config.py:13)."""
    import torch

    m = load_reference_model()
    conf = m.ConfEncoder(in_channels=80, channels=channels, n_embeddings=512, z_dim=64, c_dim=256)
    torch.manual_seed(seed)
    enc = m.Encoder(conf)
    enc.eval()
    return enc


def load_reference_preprocess_mulaw():
    """Return (mulaw_encode, mulaw_decode) from the reference ``preprocess.py:20-35``.

    ``preprocess.py`` imports librosa (absent), so only the two pure-numpy functions are
    exec'd, located by their ``def`` lines.
    """
    import numpy as np

    path = os.path.join(REFERENCE_ROOT, "preprocess.py")
    with open(path, "r") as f:
        src = f.read().splitlines(keepends=True)
    start = next(i for i, l in enumerate(src) if l.startswith("def mulaw_encode"))
    end = next(i for i, l in enumerate(src) if l.startswith("@dataclass") and i > start)
    ns = {"np": np, "ND_FP32": object, "ND_LONG": object}
    exec(compile("".join(src[start:end]), path, "exec"), ns)
    return ns["mulaw_encode"], ns["mulaw_decode"]
