"""Checkpoint ingestion for upstream-format weights.

The reference keeps its weights in three container layouts; all of them end in plain ``state_dict``s whose keys this
package's ``Encoder`` / ``Vocoder`` accept unchanged:

* ``{"encoder": sd, "cpc": ..., "optimizer": ..., "scheduler": ..., "epoch": n}`` -- ``train_cpc.py:23-29``
  (``save_checkpoint``), read back by ``convert.py:39`` and ``encode.py:29`` as ``checkpoint["encoder"]``;
* ``{"vocoder": sd, ...}`` -- the upstream release format read by ``convert.py:44`` as ``checkpoint["vocoder"]``;
* a PyTorch-Lightning checkpoint ``{"state_dict": {...}}`` of ``VocoderModel`` (``vocoder.py:41-51``): the vocoder
  lives under the ``model.`` prefix, the frozen encoder under ``encoder.``.

Nothing here touches the GPU; ``load_encoder`` / ``load_vocoder`` return modules on the CPU (call ``.to("cuda")``).
"""
from __future__ import annotations

import os
from typing import Any, Dict, Mapping, Optional, Union

import torch
from torch import Tensor

from .model import ConfEncoder, Encoder
from .network_vocoder import ConfVocoder, Vocoder

StateDict = Dict[str, Tensor]
_ENCODER_ROOTS = ("conv.", "encoder.", "codebook.", "rnn.")


def _read(ckpt: Union[str, os.PathLike, Mapping[str, Any]], trust_pickle: bool = False) -> Mapping[str, Any]:
    if isinstance(ckpt, Mapping):
        return ckpt
    # Same call as the reference (convert.py:38) -- tensors stay on the CPU -- except that the safe tensor-only unpickler is
    # the default: release checkpoints are downloaded files and only plain tensor state_dicts are consumed here.  A
    # Lightning checkpoint that pickles hyper-parameter objects needs ``trust_pickle=True`` (runs arbitrary pickled code).
    try:
        return torch.load(os.fspath(ckpt), map_location="cpu", weights_only=not trust_pickle)
    except Exception as e:  # noqa: BLE001 -- re-raised with the remedy spelled out
        if trust_pickle:
            raise
        raise RuntimeError(f"{os.fspath(ckpt)}: not loadable with the tensor-only unpickler ({type(e).__name__}: {e}).  If the "
                           "file is trusted (e.g. a Lightning checkpoint carrying non-tensor hyper-parameters), pass "
                           "trust_pickle=True.") from e


def _strip(sd: Mapping[str, Tensor], prefix: str) -> StateDict:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def extract_state_dicts(ckpt: Union[str, os.PathLike, Mapping[str, Any]], trust_pickle: bool = False) -> Dict[str, StateDict]:
    """Returns ``{"encoder": sd}`` and/or ``{"vocoder": sd}`` -- whatever the container holds.

    Raises ``KeyError`` if neither is found (the message lists the top-level keys seen)."""
    c = _read(ckpt, trust_pickle)
    out: Dict[str, StateDict] = {}
    if isinstance(c.get("encoder"), Mapping):
        out["encoder"] = dict(c["encoder"])
    if isinstance(c.get("vocoder"), Mapping):
        out["vocoder"] = dict(c["vocoder"])
    sd = c.get("state_dict")
    if isinstance(sd, Mapping):                       # Lightning: VocoderModel.model / VocoderModel.encoder
        voc = _strip(sd, "model.")
        if voc and "vocoder" not in out:
            out["vocoder"] = voc
        enc = _strip(sd, "encoder.")
        # VocoderModel.encoder is the whole Encoder, whose own children are conv / encoder / codebook / rnn
        if enc and all(k.startswith(_ENCODER_ROOTS) for k in enc) and "encoder" not in out:
            out["encoder"] = enc
    if not out and c and all(isinstance(v, Tensor) for v in c.values()):
        keys = list(c.keys())                         # a bare state_dict
        if all(k.startswith(_ENCODER_ROOTS) for k in keys):
            out["encoder"] = dict(c)
        elif any(k.startswith(("code_embedding.", "speaker_embedding.", "rnnms.", "model.")) for k in keys):
            out["vocoder"] = dict(c)
    if not out:
        raise KeyError(f"no encoder / vocoder weights found; top-level keys: {sorted(map(str, c.keys()))[:12]}")
    return out


def encoder_conf_from_state(sd: Mapping[str, Tensor]) -> ConfEncoder:
    """Recovers the constructor fields from tensor shapes (``model.py:43-57``)."""
    conv = sd["conv.weight"]                           # (channels, in_channels, 4)
    emb = sd["codebook.embedding"]                     # (n_embeddings, z_dim)
    c_dim = sd["rnn.weight_hh_l0"].shape[1]            # (4 c_dim, c_dim)
    return ConfEncoder(in_channels=conv.shape[1], channels=conv.shape[0], n_embeddings=emb.shape[0], z_dim=emb.shape[1],
                       c_dim=c_dim)


def vocoder_conf_from_state(sd: Mapping[str, Tensor]) -> ConfVocoder:
    """Recovers ``size_i_codebook``, ``dim_i_embedding``, ``n_speakers`` and ``dim_speaker_embedding`` from the two
    embedding tables (``network_vocoder.py:36-38``); the RNNMS core keeps its fixed dimensions (``config.py:62-77``)."""
    sd = Vocoder.remap_state_dict(dict(sd))
    code, spk = sd["code_embedding.weight"], sd["speaker_embedding.weight"]
    return ConfVocoder(size_i_codebook=code.shape[0], dim_i_embedding=code.shape[1], n_speakers=spk.shape[0],
                       dim_speaker_embedding=spk.shape[1])


def load_encoder(ckpt: Union[str, os.PathLike, Mapping[str, Any]], conf: Optional[ConfEncoder] = None,
                 trust_pickle: bool = False) -> Encoder:
    """``Encoder`` in eval mode with the checkpoint's weights (``convert.py:32,39,47`` in one call)."""
    sd = extract_state_dicts(ckpt, trust_pickle)["encoder"]
    enc = Encoder(conf if conf is not None else encoder_conf_from_state(sd))
    enc.load_state_dict(sd)
    return enc.eval()


def load_vocoder(ckpt: Union[str, os.PathLike, Mapping[str, Any]], conf: Optional[ConfVocoder] = None,
                 trust_pickle: bool = False) -> Vocoder:
    """``Vocoder`` in eval mode with the checkpoint's weights (``convert.py:33,44,48`` in one call)."""
    sd = extract_state_dicts(ckpt, trust_pickle)["vocoder"]
    voc = Vocoder(conf if conf is not None else vocoder_conf_from_state(sd))
    voc.load_state_dict(sd)
    return voc.eval()
