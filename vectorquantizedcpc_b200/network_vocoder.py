"""Host-side mirror of /root/reference/network_vocoder.py:11-78 (``ConfVocoder``, ``Vocoder``) plus the
``rnnms`` vocoder core it delegates to (``RNNMSVocoder`` of tarepan/UniversalVocoding -- NOT in the reference
tree, restated per SURVEY.md App. A.3 with the dimensions pinned by /root/reference/config.py:62-77,199).

``Vocoder.generate(z, speaker) -> wav (B, L)`` and ``Vocoder.forward(x, z, speaker) -> energies (B, L, 256)``
keep the reference signatures; ``generate`` adds optional, non-breaking keywords for injected randomness.
The ``nn`` sub-modules are parameter containers only; all arithmetic runs in sm_100a CUDA (include/vqcpc.h).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib
from .model import _check_no_grad
from .mulaw import mulaw_decode_table


@dataclass
class ConfPrenet:
    num_layers: int = 2            # config.py:72
    bidirectional: bool = True     # config.py:73


@dataclass
class ConfWaveAR:
    size_i_embed_ar: int = 256     # config.py:75
    size_h_rnn: int = 896          # config.py:76
    size_h_fc: int = 256           # config.py:77


@dataclass
class ConfRNNMSVocoder:
    """Config surface of rnnms' vocoder as pinned in-tree by /root/reference/config.py:67-77,199."""
    dim_i_feature: int = 128       # dim_i_embedding + dim_speaker_embedding, config.py:199
    dim_voc_latent: int = 256      # config.py:68
    bits_mu_law: int = 8           # config.py:69
    upsampling_t: int = 160        # config.py:70 -> hop_length, config.py:102
    prenet: ConfPrenet = field(default_factory=ConfPrenet)
    wave_ar: ConfWaveAR = field(default_factory=ConfWaveAR)


@dataclass
class ConfVocoder:
    """/root/reference/network_vocoder.py:11-24."""
    size_i_codebook: int = 512
    dim_i_embedding: int = 64
    n_speakers: int = 102
    dim_speaker_embedding: int = 64
    rnnms: ConfRNNMSVocoder = field(default_factory=ConfRNNMSVocoder)


class _PreNet(nn.Module):
    def __init__(self, dim_i: int, dim_o: int, num_layers: int, bidirectional: bool):
        super().__init__()
        self.net = nn.GRU(dim_i, dim_o // (2 if bidirectional else 1), num_layers=num_layers, batch_first=True,
                          bidirectional=bidirectional)


class _WaveAR(nn.Module):
    def __init__(self, dim_cond: int, conf: ConfWaveAR, n_classes: int):
        super().__init__()
        self.embedding = nn.Embedding(n_classes, conf.size_i_embed_ar)
        self.rnn = nn.GRU(conf.size_i_embed_ar + dim_cond, conf.size_h_rnn, batch_first=True)
        self.fc1 = nn.Linear(conf.size_h_rnn, conf.size_h_fc)
        self.fc2 = nn.Linear(conf.size_h_fc, n_classes)


class RNNMSVocoder(nn.Module):
    """Parameter container for the rnnms core (prenet biGRU + autoregressive GRU/fc1/fc2).  Sub-module names
    are ours (upstream names are not pinned in the reference tree); see ``REMAP`` in ``Vocoder``."""

    def __init__(self, conf: ConfRNNMSVocoder):
        super().__init__()
        self.conf = conf
        self.prenet = _PreNet(conf.dim_i_feature, conf.dim_voc_latent, conf.prenet.num_layers, conf.prenet.bidirectional)
        self.ar = _WaveAR(conf.dim_voc_latent, conf.wave_ar, 1 << conf.bits_mu_law)


# The persistent kernels are specialised for the dimensions the reference pins (config.py:62-77,199).
_SUPPORTED = dict(dim_i_feature=128, dim_voc_latent=256, bits_mu_law=8, num_layers=2, bidirectional=True,
                  size_i_embed_ar=256, size_h_rnn=896, size_h_fc=256)


class Vocoder(nn.Module):
    """/root/reference/network_vocoder.py:26-78: bidirectional PreNet + WaveRNN (= RNN_MS) conditioned on
    discrete VQ-CPC codes and a speaker id."""

    # alternative checkpoint key prefixes -> ours (bshall/UniversalVocoding-style names and Lightning's "model.")
    REMAP = (("rnnms.rnn1.", "rnnms.prenet.net."), ("rnnms.embedding.", "rnnms.ar.embedding."),
             ("rnnms.rnn2.", "rnnms.ar.rnn."), ("rnnms.fc1.", "rnnms.ar.fc1."), ("rnnms.fc2.", "rnnms.ar.fc2."),
             ("rnn1.", "rnnms.prenet.net."), ("embedding.", "rnnms.ar.embedding."), ("rnn2.", "rnnms.ar.rnn."),
             ("fc1.", "rnnms.ar.fc1."), ("fc2.", "rnnms.ar.fc2."))

    def __init__(self, conf: ConfVocoder | None = None, **kwargs):
        super().__init__()
        if conf is None:
            conf = ConfVocoder(**kwargs)
        elif kwargs:
            raise TypeError("pass either a ConfVocoder or keyword fields, not both")
        r = conf.rnnms
        got = dict(dim_i_feature=r.dim_i_feature, dim_voc_latent=r.dim_voc_latent, bits_mu_law=r.bits_mu_law,
                   num_layers=r.prenet.num_layers, bidirectional=r.prenet.bidirectional,
                   size_i_embed_ar=r.wave_ar.size_i_embed_ar, size_h_rnn=r.wave_ar.size_h_rnn,
                   size_h_fc=r.wave_ar.size_h_fc)
        if got != _SUPPORTED:
            raise ValueError(f"the sm_100a kernels are specialised for {_SUPPORTED}; got {got}")
        if conf.dim_i_embedding + conf.dim_speaker_embedding != r.dim_i_feature:
            raise ValueError("dim_i_embedding + dim_speaker_embedding must equal rnnms.dim_i_feature (config.py:199)")
        self.conf = conf
        self.code_embedding = nn.Embedding(conf.size_i_codebook, conf.dim_i_embedding)
        self.speaker_embedding = nn.Embedding(conf.n_speakers, conf.dim_speaker_embedding)
        self.rnnms = RNNMSVocoder(conf.rnnms)
        self._packed = None
        self._packed_key = None

    # -------------------------------------------------------------------------------- checkpoints
    @classmethod
    def remap_state_dict(cls, sd: Dict[str, Tensor]) -> Dict[str, Tensor]:
        """Accept Lightning ``model.``-prefixed checkpoints (vocoder.py:47) and rnn1/rnn2-style core names."""
        out = {}
        for k, v in sd.items():
            if k.startswith("model."):
                k = k[len("model."):]
            if not k.startswith(("code_embedding.", "speaker_embedding.", "rnnms.prenet.", "rnnms.ar.")):
                for old, new in cls.REMAP:
                    if k.startswith(old):
                        k = new + k[len(old):]
                        break
            out[k] = v
        return out

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        return super().load_state_dict(self.remap_state_dict(state_dict), strict=strict, **kw)

    # -------------------------------------------------------------------------------- weights
    def _weight_tensors(self):
        g, ar = self.rnnms.prenet.net, self.rnnms.ar
        ts = [self.code_embedding.weight, self.speaker_embedding.weight]
        for layer in (0, 1):
            for name in ("weight_ih", "bias_ih", "weight_hh", "bias_hh"):
                ts += [getattr(g, f"{name}_l{layer}"), getattr(g, f"{name}_l{layer}_reverse")]
        ts += [ar.rnn.weight_ih_l0, ar.rnn.bias_ih_l0, ar.rnn.weight_hh_l0, ar.rnn.bias_hh_l0,
               ar.fc1.weight, ar.fc1.bias, ar.fc2.weight, ar.fc2.bias, ar.embedding.weight]
        return ts

    def pack_weights(self):
        """One-time re-layout for the kernels (cached on parameter storage + version): directions of each
        prenet layer concatenated, E' = emb . W_ih[:, :256]^T precomputed on the device, mu-law table."""
        ts = self._weight_tensors()
        key = tuple((t.data_ptr(), t._version, t.device) for t in ts)
        if self._packed is not None and key == self._packed_key:
            return self._packed
        for t in ts:
            _lib.require_cuda(t, "Vocoder parameter")
        f = lambda t: t.detach().to(torch.float32).contiguous()
        g, ar = self.rnnms.prenet.net, self.rnnms.ar
        dev = self.code_embedding.weight.device
        keep = []
        w = _lib.VocoderWeights()
        w.n_codes, w.dim_code = self.conf.size_i_codebook, self.conf.dim_i_embedding
        w.n_speakers, w.dim_speaker = self.conf.n_speakers, self.conf.dim_speaker_embedding
        w.upsample_t = self.conf.rnnms.upsampling_t

        def put(name, t, index=None):
            t = f(t)
            keep.append(t)
            if index is None:
                setattr(w, name, t.data_ptr())
            else:
                getattr(w, name)[index] = t.data_ptr()
            return t

        put("code_emb", self.code_embedding.weight)
        put("spk_emb", self.speaker_embedding.weight)
        for layer in (0, 1):
            cat = lambda n: torch.cat([getattr(g, f"{n}_l{layer}").detach(), getattr(g, f"{n}_l{layer}_reverse").detach()], 0)
            put("pre_w_ih", cat("weight_ih"), layer)
            put("pre_b_ih", cat("bias_ih"), layer)
            put("pre_w_hh", cat("weight_hh"), layer)
            put("pre_b_hh", cat("bias_hh"), layer)
        put("ar_w_ih", ar.rnn.weight_ih_l0)
        put("ar_b_ih", ar.rnn.bias_ih_l0)
        put("ar_w_hh", ar.rnn.weight_hh_l0)
        put("ar_b_hh", ar.rnn.bias_hh_l0)
        put("fc1_w", ar.fc1.weight)
        put("fc1_b", ar.fc1.bias)
        put("fc2_w", ar.fc2.weight)
        put("fc2_b", ar.fc2.bias)
        put("ar_emb", ar.embedding.weight)
        put("mulaw_lut", torch.from_numpy(mulaw_decode_table(self.conf.rnnms.bits_mu_law)).to(dev))
        eprime = torch.empty(256, 3 * 896, device=dev)
        keep.append(eprime)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().vqcpc_vocoder_pack(C.byref(w), _lib.ptr(eprime), _lib.current_stream_ptr()),
                       "Vocoder.pack_weights")
        w.eprime = eprime.data_ptr()
        self._packed, self._packed_key = (w, keep), key
        return self._packed

    # -------------------------------------------------------------------------------- hot path
    BATCHED_MIN_B = 8        # from this many utterances a launch runs the batched tensor-core kernels (csrc/vocoder.cu)
    BATCHED_GROUP = 128      # utterances per launch of the two-group batched kernel

    def _check_inputs(self, z: Tensor, speaker: Tensor):
        _lib.require_cuda(z, "z")
        _lib.require_cuda(speaker, "speaker")
        if z.dim() != 2 or speaker.dim() != 1 or speaker.shape[0] != z.shape[0]:
            raise ValueError(f"z must be (B, Tc) and speaker (B,), got {tuple(z.shape)} and {tuple(speaker.shape)}")
        if z.dtype != torch.int64 or speaker.dtype != torch.int64:
            raise ValueError("z and speaker must be int64 (LongTensor), as torch.argmin / convert.py:73 produce")
        if z.shape[1] < 1:
            raise ValueError("z needs at least one code frame")
        # the index range (nn.Embedding raises IndexError) is checked ON THE DEVICE by the gather kernel and surfaces through the
        # workspace status word at the call's single synchronisation point -- no extra host sync before the launch

    def _check_lengths(self, lengths, z: Tensor) -> Tensor:
        """``lengths`` (B,) = valid code frames per utterance of a padded batch -> int32 device tensor."""
        lengths = torch.as_tensor(lengths)
        if lengths.dim() != 1 or lengths.shape[0] != z.shape[0] or lengths.dtype.is_floating_point:
            raise ValueError("lengths must be an integer vector with one entry per utterance")
        if lengths.numel() and (int(lengths.min()) < 1 or int(lengths.max()) > z.shape[1]):
            raise ValueError(f"lengths must lie in [1, {z.shape[1]}]")
        return lengths.to(device=z.device, dtype=torch.int32).contiguous()

    def _condition(self, z: Tensor, speaker: Tensor, lengths, ws: Tensor, return_prenet: bool = False):
        """Launches the conditioning path on workspace ``ws`` (>= vqcpc_vocoder_workspace_bytes(B, Tc)); no status check."""
        w, _keep = self.pack_weights()
        B, Tc = z.shape
        dev = z.device
        lib = _lib.lib()
        G = torch.empty(B, 2 * Tc, 3 * 896, device=dev)
        p = torch.empty(B, 2 * Tc, 256, device=dev) if return_prenet else None
        zc, sc = z.contiguous(), speaker.contiguous()
        with torch.cuda.device(dev):
            if lengths is None:
                st = lib.vqcpc_vocoder_condition(C.byref(w), _lib.ptr(zc), _lib.ptr(sc), B, Tc, _lib.ptr(ws), ws.numel(),
                                                 _lib.ptr(G), _lib.ptr(p), _lib.current_stream_ptr())
            else:
                st = lib.vqcpc_vocoder_condition_ragged(C.byref(w), _lib.ptr(zc), _lib.ptr(sc), _lib.ptr(lengths), B, Tc,
                                                        _lib.ptr(ws), ws.numel(), _lib.ptr(G), _lib.ptr(p),
                                                        _lib.current_stream_ptr())
            _lib.check(st, "Vocoder.condition")
        return G, p

    def _workspace(self, z: Tensor) -> Tensor:
        B, Tc = z.shape
        return torch.empty(_lib.lib().vqcpc_vocoder_workspace_bytes(B, Tc), dtype=torch.uint8, device=z.device)

    def condition(self, z: Tensor, speaker: Tensor, return_prenet: bool = False, lengths=None):
        """Embeddings + x2 nearest + concat (network_vocoder.py:73-77), prenet biGRU, hoisted input projection
        G (B, 2Tc, 2688).  With ``return_prenet`` also the prenet output p (B, 2Tc, 256).  ``lengths`` (B,): ragged
        batch -- utterance b's bidirectional prenet runs over its own 2*lengths[b] frames, the padded tail is zero."""
        self._check_inputs(z, speaker)
        if lengths is not None:
            lengths = self._check_lengths(lengths, z)
        ws = self._workspace(z)
        G, p = self._condition(z, speaker, lengths, ws, return_prenet)
        if z.numel():
            with torch.cuda.device(z.device):
                _lib.check(_lib.lib().vqcpc_check_status(_lib.ptr(ws), _lib.current_stream_ptr()), "Vocoder.condition")
        return (G, p) if return_prenet else G

    def generate(self, z: Tensor, speaker: Tensor, uniforms: Optional[Tensor] = None, return_mulaw: bool = False,
                 n_steps: Optional[int] = None, generator: Optional[torch.Generator] = None,
                 return_logits: bool = False, lengths=None):
        """Generate utterances from a batch of (latent_code, speaker_index) -- network_vocoder.py:69-78.

        z (B, Tc) int64, speaker (B,) int64 -> wav (B, L) fp32 on the input device, L = 320*Tc (or ``n_steps``).
        ``uniforms`` (B, L) in [0,1) injects the sampler's randomness (one per utterance and step); when None
        they are drawn with ``torch.rand(generator=generator)`` on the device.  One persistent kernel launch per
        group of utterances; no per-sample host work.

        ``lengths`` (B,) int: ragged batch -- z is padded to the longest utterance and utterance b
        has lengths[b] valid code frames.  Its first 320*lengths[b] samples are exactly what an unpadded call with
        the same uniforms returns; the rest of its row is set to 0 (use ``wav[b, :320*lengths[b]]``)."""
        _check_no_grad()
        self._check_inputs(z, speaker)
        len_dev = self._check_lengths(lengths, z) if lengths is not None else None
        ws = self._workspace(z)               # shared by the conditioning path and the sample loop: ONE status word, one sync
        G, _ = self._condition(z, speaker, len_dev, ws)
        w, _keep = self.pack_weights()
        B, Tc = z.shape
        dev = z.device
        L = 2 * Tc * self.conf.rnnms.upsampling_t if n_steps is None else int(n_steps)
        if L < 0 or L > 2 * Tc * self.conf.rnnms.upsampling_t:
            raise ValueError(f"n_steps={L} outside [0, {2 * Tc * self.conf.rnnms.upsampling_t}]")
        if uniforms is None:
            uniforms = torch.rand(B, L, device=dev, generator=generator)
        else:
            _lib.require_cuda(uniforms, "uniforms")
            if tuple(uniforms.shape) != (B, L):
                raise ValueError(f"uniforms must be ({B}, {L}), got {tuple(uniforms.shape)}")
            uniforms = uniforms.to(torch.float32).contiguous()
        lib = _lib.lib()
        up2 = 2 * self.conf.rnnms.upsampling_t
        if lengths is not None and B > 1 and L > 0:
            # Ragged batch: length-sorted launch groups, each run only as far as ITS longest utterance (per-utterance stop for
            # the single-utterance kernel, buckets of the batched kernels' natural group size otherwise) -- vocoder.py:69
            # "cannot batch" / SURVEY 8f row 2.  The sample loop is causal, so an utterance's valid prefix does not depend on
            # how far its launch runs; samples beyond 320*lengths[b] are 0.
            len_host = torch.as_tensor(lengths).to("cpu", torch.int64)
            steps = torch.clamp(len_host * up2, max=L)
            order = torch.argsort(steps, descending=True, stable=True)
            group = 1 if B < self.BATCHED_MIN_B else self.BATCHED_GROUP
            wav = torch.zeros(B, L, device=dev)
            codes = torch.zeros(B, L, dtype=torch.int32, device=dev) if return_mulaw else None
            logits = torch.zeros(B, L, 256, device=dev) if return_logits else None
            with torch.cuda.device(dev):
                for g0 in range(0, B, group):
                    sel = order[g0:g0 + group]
                    Lg = int(steps[sel[0]])                      # the group's longest utterance
                    if Lg == 0:
                        continue
                    seld = sel.to(dev)
                    nb = int(sel.numel())
                    whole = nb == B and Lg == L
                    Gg = G if whole else G.index_select(0, seld)
                    ug = uniforms if whole else uniforms.index_select(0, seld)[:, :Lg].contiguous()
                    wg = torch.empty(nb, Lg, device=dev)
                    cg = torch.empty(nb, Lg, dtype=torch.int32, device=dev) if return_mulaw else None
                    lg = torch.empty(nb, Lg, 256, device=dev) if return_logits else None
                    st = lib.vqcpc_vocoder_generate(C.byref(w), _lib.ptr(Gg), _lib.ptr(ug), nb, 2 * Tc, Lg, _lib.ptr(ws), ws.numel(),
                                                    _lib.ptr(wg), _lib.ptr(cg), _lib.ptr(lg), _lib.current_stream_ptr())
                    _lib.check(st, "Vocoder.generate")
                    # every launch resets the workspace status word: read it before the next one (a launch is milliseconds)
                    _lib.check(lib.vqcpc_check_status(_lib.ptr(ws), _lib.current_stream_ptr()), "Vocoder.generate")
                    valid = torch.arange(Lg, device=dev)[None, :] < steps[sel].to(dev)[:, None]
                    wav[seld, :Lg] = wg * valid
                    if return_mulaw:
                        codes[seld, :Lg] = cg * valid
                    if return_logits:
                        logits[seld, :Lg] = lg * valid[..., None]
        else:
            wav = torch.empty(B, L, device=dev)
            codes = torch.empty(B, L, dtype=torch.int32, device=dev) if return_mulaw else None
            logits = torch.empty(B, L, 256, device=dev) if return_logits else None
            with torch.cuda.device(dev):
                st = lib.vqcpc_vocoder_generate(C.byref(w), _lib.ptr(G), _lib.ptr(uniforms), B, 2 * Tc, L, _lib.ptr(ws),
                                                ws.numel(), _lib.ptr(wav), _lib.ptr(codes), _lib.ptr(logits),
                                                _lib.current_stream_ptr())
                _lib.check(st, "Vocoder.generate")
                if B > 0:
                    _lib.check(lib.vqcpc_check_status(_lib.ptr(ws), _lib.current_stream_ptr()), "Vocoder.generate")
            if lengths is not None and B > 0 and L > 0:
                valid = (torch.arange(L, device=dev)[None, :] <
                         (torch.as_tensor(lengths, device=dev).to(torch.int64) * up2)[:, None])
                wav = wav * valid
        out = (wav,)
        if return_mulaw:
            out += (codes.to(torch.int64),)
        if return_logits:
            out += (logits,)
        return out[0] if len(out) == 1 else out

    def forward(self, x: Tensor, z: Tensor, speaker: Tensor, lengths=None) -> Tensor:
        """Teacher-forced energies (B, L, 256) -- network_vocoder.py:41-67; x (B, L) int64 mu-law series is the
        AR input (vocoder.py:62: ``audio_series[:, :-1]``).  Inference-only (no autograd graph).  ``lengths``: as in
        ``generate`` (energies beyond 320*lengths[b] belong to the padding and carry no meaning)."""
        _check_no_grad()
        _lib.require_cuda(x, "x")
        if x.dim() != 2 or x.shape[0] != z.shape[0] or x.dtype != torch.int64:
            raise ValueError("x must be (B, L) int64")
        if x.numel() and bool(((x < 0) | (x > 255)).any()):
            raise IndexError("mu-law code out of range [0, 255]")
        self._check_inputs(z, speaker)
        len_dev = self._check_lengths(lengths, z) if lengths is not None else None
        ws = self._workspace(z)
        G, _ = self._condition(z, speaker, len_dev, ws)
        w, _keep = self.pack_weights()
        B, Tc = z.shape
        L = x.shape[1]
        if L > 2 * Tc * self.conf.rnnms.upsampling_t:
            raise ValueError(f"x is longer ({L}) than the conditioning series ({2 * Tc * self.conf.rnnms.upsampling_t})")
        dev = z.device
        lib = _lib.lib()
        logits = torch.empty(B, L, 256, device=dev)
        xc = x.contiguous()
        with torch.cuda.device(dev):
            st = lib.vqcpc_vocoder_logits_tf(C.byref(w), _lib.ptr(G), _lib.ptr(xc), B, 2 * Tc, L, _lib.ptr(ws), ws.numel(),
                                             _lib.ptr(logits), _lib.current_stream_ptr())
            _lib.check(st, "Vocoder.forward")
            if B > 0:
                _lib.check(lib.vqcpc_check_status(_lib.ptr(ws), _lib.current_stream_ptr()), "Vocoder.forward")
        return logits
