"""Host-side mirror of /root/reference/model.py:17-155 (``ConfEncoder``, ``Encoder``, ``VQEmbeddingEMA``).

Same class names, constructor fields, attribute names (``conv``, ``encoder``, ``codebook``, ``rnn``) and
``state_dict`` layout as the reference, so ``load_state_dict`` of a reference checkpoint and the forward
hook of encode.py:34-40 resolve unchanged.  The ``nn`` sub-modules are PARAMETER CONTAINERS only: every
operation of ``encode`` runs in hand-written sm_100a CUDA behind the C ABI of include/vqcpc.h.
There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from itertools import chain
from typing import Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib


@dataclass
class ConfEncoder:
    """/root/reference/model.py:17-31 (defaults = the values config.py:28-33 resolves to, channels per the
    north-star headline; the reference YAML default is 512)."""
    in_channels: int = 80
    channels: int = 768
    n_embeddings: int = 512
    z_dim: int = 64
    c_dim: int = 256


def _check_no_grad(*tensors: Tensor) -> None:
    if torch.is_grad_enabled() and any(t.requires_grad for t in tensors):
        raise RuntimeError("vectorquantizedcpc_b200 is inference-only: call under torch.no_grad() "
                           "(as convert.py:75 / encode.py:45 do) or detach the inputs")


class VQEmbeddingEMA(nn.Module):
    """/root/reference/model.py:89-115.  ``encode`` is the hot path; ``forward`` (EMA codebook update,
    commitment loss -- model.py:117-155) is training-only and out of scope."""

    def __init__(self, n_embeddings: int, embedding_dim: int, commitment_cost: float = 0.25, decay: float = 0.999,
                 epsilon: float = 1e-5):
        super().__init__()
        self.commitment_cost = commitment_cost
        self.decay = decay
        self.epsilon = epsilon
        init_bound = 1 / 512
        embedding = torch.Tensor(n_embeddings, embedding_dim)
        embedding.uniform_(-init_bound, init_bound)
        self.register_buffer("embedding", embedding)
        self.register_buffer("ema_count", torch.zeros(n_embeddings))
        self.register_buffer("ema_weight", self.embedding.clone())

    def encode(self, x: Tensor) -> Tuple[Tensor, Tensor]:
        """x (B, T, D) -> (quantized (B, T, D), indices (B, T) int64) -- model.py:103-115."""
        _lib.require_cuda(x, "x")
        _lib.require_cuda(self.embedding, "codebook")
        _check_no_grad(x)
        if x.dim() != 3 or x.shape[-1] != self.embedding.shape[1]:
            raise ValueError(f"x must be (B, T, {self.embedding.shape[1]}), got {tuple(x.shape)}")
        xf = x.detach().to(torch.float32).contiguous()
        cb = self.embedding.detach().to(torch.float32).contiguous()
        B, T, D = xf.shape
        q = torch.empty_like(xf)
        idx = torch.empty(B, T, dtype=torch.int64, device=xf.device)
        lib = _lib.lib()
        with torch.cuda.device(xf.device):
            tc = B * T >= 8192      # tensor-core pipeline: caller-owned workspace (codebook planes + status word)
            ws_bytes = lib.vqcpc_vq_workspace_bytes() if tc else 0
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xf.device) if tc else None
            st = lib.vqcpc_vq_lookup(_lib.ptr(xf), _lib.ptr(cb), B * T, cb.shape[0], D, _lib.ptr(q), _lib.ptr(idx),
                                     _lib.ptr(ws), ws_bytes, _lib.current_stream_ptr())
            _lib.check(st, "VQEmbeddingEMA.encode")
            if tc:                  # surface a device-side timeout instead of returning garbage
                _lib.check(lib.vqcpc_check_status(_lib.ptr(ws), _lib.current_stream_ptr()), "VQEmbeddingEMA.encode")
        return q, idx

    def forward(self, x):
        raise NotImplementedError("VQEmbeddingEMA.forward is the training path (EMA update, model.py:117-155): "
                                  "out of scope of the inference hot path; use .encode()")


def _repeat_modules(atom_gen, n_repeat):
    return list(chain.from_iterable([atom_gen() for _ in range(n_repeat)]))


class Encoder(nn.Module):
    """/root/reference/model.py:33-70.  Conv1d/k4s2 - LN - ReLU - [FC - LN - ReLU]x4 - FC - VQ + LSTM."""

    def __init__(self, conf: ConfEncoder | None = None, **kwargs):
        super().__init__()
        if conf is None:
            conf = ConfEncoder(**kwargs)        # stale call style Encoder(**cfg.model.encoder), convert.py:32
        elif kwargs:
            raise TypeError("pass either a ConfEncoder or keyword fields, not both")
        self.conf = conf
        # construction order == reference (same RNG stream under the same seed)
        self.conv = nn.Conv1d(conf.in_channels, conf.channels, 4, 2, 1, bias=False)
        self.encoder = nn.Sequential(
            nn.LayerNorm(conf.channels),
            nn.ReLU(True),
            *_repeat_modules(lambda: [nn.Linear(conf.channels, conf.channels, bias=False),
                                      nn.LayerNorm(conf.channels), nn.ReLU(True)], 4),
            nn.Linear(conf.channels, conf.z_dim),
        )
        self.codebook = VQEmbeddingEMA(conf.n_embeddings, conf.z_dim)
        self.rnn = nn.LSTM(conf.z_dim, conf.c_dim, batch_first=True)
        self._packed = None
        self._packed_key = None
        # GEMM arithmetic: "fp32" = CUDA-core FMA (exact-order parity path), "bf16x3" = tcgen05 tensor cores with a
        # bf16 hi/lo split (fp32-grade: ~2^-16 relative per product), "auto" = bf16x3 once B*T' >= AUTO_TC_ROWS (so the
        # arithmetic -- and, for near-tie frames, an index -- can depend on the batch size; set an explicit mode for results
        # that do not), "bf16" = speed mode: single-pass bf16 products in the conv / MLP / projection (z_pre within 3e-2 of
        # its scale, >= 97 % index agreement on trained-like weights; VQ exact for that z, LSTM still bf16x3).
        self.gemm_mode = "auto"

    # measured on B200 (3 s utterances, ms per encode, fp32 / bf16x3): 8 utterances 0.77 / 0.70, 12: 1.05 / 0.66, 27: 1.52 / 1.10
    AUTO_TC_ROWS = 1536

    def _resolve_mode(self, rows: int) -> int:
        mode = self.gemm_mode
        if mode == "auto":
            mode = "bf16x3" if rows >= self.AUTO_TC_ROWS else "fp32"
        if mode not in ("fp32", "bf16x3", "bf16"):
            raise ValueError(f"gemm_mode must be 'auto', 'fp32', 'bf16x3' or 'bf16', got {self.gemm_mode!r}")
        return {"fp32": _lib.GEMM_FP32, "bf16x3": _lib.GEMM_BF16X3, "bf16": _lib.GEMM_BF16}[mode]

    # -------------------------------------------------------------------------------- weights
    def _weight_tensors(self):
        e = self.encoder
        return [self.conv.weight, *[e[i].weight for i in (0, 3, 6, 9, 12)], *[e[i].bias for i in (0, 3, 6, 9, 12)],
                *[e[i].weight for i in (2, 5, 8, 11)], e[14].weight, e[14].bias, self.codebook.embedding,
                self.rnn.weight_ih_l0, self.rnn.weight_hh_l0, self.rnn.bias_ih_l0, self.rnn.bias_hh_l0]

    def pack_weights(self):
        """Build (and cache, keyed on parameter storage + version) the ``vqcpc_encoder_weights`` struct."""
        ts = self._weight_tensors()
        key = tuple((t.data_ptr(), t._version, t.device) for t in ts)
        if self._packed is not None and key == self._packed_key:
            return self._packed
        for t in ts:
            _lib.require_cuda(t, "Encoder parameter")
        keep = [t.detach().to(torch.float32).contiguous() for t in ts]
        (conv_w, l0, l1, l2, l3, l4, b0, b1, b2, b3, b4, f0, f1, f2, f3, pw, pb, cb, wih, whh, bih, bhh) = keep
        lstm_b = (bih + bhh).contiguous()
        keep.append(lstm_b)
        w = _lib.EncoderWeights()
        w.in_channels, w.channels = self.conf.in_channels, self.conf.channels
        w.n_embeddings, w.z_dim, w.c_dim = self.conf.n_embeddings, self.conf.z_dim, self.conf.c_dim
        w.conv_w = conv_w.data_ptr()
        for i, t in enumerate((l0, l1, l2, l3, l4)):
            w.ln_w[i] = t.data_ptr()
        for i, t in enumerate((b0, b1, b2, b3, b4)):
            w.ln_b[i] = t.data_ptr()
        for i, t in enumerate((f0, f1, f2, f3)):
            w.fc_w[i] = t.data_ptr()
        w.proj_w, w.proj_b, w.codebook = pw.data_ptr(), pb.data_ptr(), cb.data_ptr()
        w.lstm_w_ih, w.lstm_w_hh, w.lstm_b = wih.data_ptr(), whh.data_ptr(), lstm_b.data_ptr()
        # bf16 hi/lo planes of the GEMM weights for the tensor-core mode
        lib = _lib.lib()
        dev = conv_w.device

        def planes(t2d):
            rows, K = t2d.shape
            out = torch.empty(rows, 2 * K, dtype=torch.bfloat16, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.vqcpc_split_planes(t2d.data_ptr(), K, out.data_ptr(), rows, K, _lib.current_stream_ptr()),
                           "Encoder.pack_weights")
            keep.append(out)
            return out.data_ptr()

        w.conv_wp = planes(conv_w.view(conv_w.shape[0], -1))
        for i, t in enumerate((f0, f1, f2, f3)):
            w.fc_wp[i] = planes(t)
        w.proj_wp = planes(pw)
        w.lstm_whh_p = planes(whh)
        # weight-only precompute: the LSTM sees quantised vectors only, so its input projection is a 512-row table
        table = torch.empty(cb.shape[0], wih.shape[0], device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.vqcpc_linear_f32(cb.data_ptr(), cb.shape[1], wih.data_ptr(), wih.shape[1], lstm_b.data_ptr(),
                                            table.data_ptr(), table.shape[1], cb.shape[0], table.shape[1], cb.shape[1],
                                            _lib.current_stream_ptr()), "Encoder.pack_weights (LSTM table)")
        keep.append(table)
        w.lstm_table = table.data_ptr()
        self._packed, self._packed_key = (w, keep), key
        return self._packed

    # -------------------------------------------------------------------------------- hot path
    def encode(self, mel: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """mel (B, 80, T) fp32 -> (z (B,T',64) quantised, c (B,T',256), indices (B,T') int64) -- model.py:59-70."""
        z, c, idx, _, _ = self._encode(mel, want_aux=bool(self.encoder[-1]._forward_hooks))
        return z, c, idx

    def encode_ragged(self, mels):
        """Ragged batch: ``mels`` = sequence of (80, T_b) tensors -> list of per-utterance
        ``(z (T'_b, 64), c (T'_b, 256), indices (T'_b,))``, T'_b = (T_b - 2)//2 + 1, from ONE batched call.

        Exact, not approximate: the strided convolution pads with zeros (model.py:43), LayerNorm / Linear / VQ act
        per frame and the LSTM is causal (model.py:57), so the valid frames of a zero-padded utterance equal its
        unpadded run."""
        mels = list(mels)
        if not mels:
            return []
        for m in mels:
            if m.dim() != 2 or m.shape[0] != self.conf.in_channels or m.shape[1] < 2:
                raise ValueError(f"every mel must be ({self.conf.in_channels}, T >= 2), got {tuple(m.shape)}")
        T = max(m.shape[1] for m in mels)
        batch = mels[0].new_zeros(len(mels), self.conf.in_channels, T)
        for b, m in enumerate(mels):
            batch[b, :, :m.shape[1]] = m
        z, c, idx = self.encode(batch)
        out = []
        for b, m in enumerate(mels):
            n = (m.shape[1] - 2) // 2 + 1
            out.append((z[b, :n], c[b, :n], idx[b, :n]))
        return out

    def encode_from_host(self, mel: Tensor, chunk_utterances: int | None = None) -> Tuple[Tensor, Tensor, Tensor]:
        """``encode`` of a batch that still sits in HOST memory (pin it for a truly asynchronous copy): the utterances go to
        the GPU in chunks on a copy stream, double buffered, and each chunk's front part (conv .. VQ, per frame) runs while
        the next chunk is in flight; the LSTM -- a T'-step latency chain whose cost hardly depends on the batch -- runs once
        over all utterances at the end.  Results equal ``encode(mel.cuda())`` (frames and utterances are independent).
        Default chunk: about 512 utterances, rounded so that a chunk's rows fill whole waves of the persistent GEMM
        (SM pairs x 256 rows) -- 512 x 3 s is 4.05 waves on 148 SMs and would pay for 5."""
        if mel.is_cuda:
            return self.encode(mel)
        _check_no_grad(mel)
        if mel.dim() != 3 or mel.shape[1] != self.conf.in_channels:
            raise ValueError(f"mel must be (B, {self.conf.in_channels}, T), got {tuple(mel.shape)}")
        if mel.shape[2] < 2:
            raise ValueError("mel needs at least 2 frames (Conv1d kernel 4, stride 2, padding 1)")
        if chunk_utterances is not None and chunk_utterances < 1:
            raise ValueError("chunk_utterances must be >= 1")
        dev = self.rnn.weight_hh_l0.device
        if dev.type != "cuda":
            raise RuntimeError("Encoder.encode_from_host: the module must live on a CUDA device (no CPU fallback)")
        w, _keep = self.pack_weights()
        mel = mel.detach().to(torch.float32).contiguous()
        B, _, T = mel.shape
        Tp = (T - 2) // 2 + 1
        lib = _lib.lib()
        z = torch.empty(B, Tp, self.conf.z_dim, device=dev)
        c = torch.empty(B, Tp, self.conf.c_dim, device=dev)
        idx = torch.empty(B, Tp, dtype=torch.int64, device=dev)
        if B == 0:
            return z, c, idx
        mode = self._resolve_mode(B * Tp)
        if chunk_utterances is None:
            rows_per_wave = max(1, torch.cuda.get_device_properties(dev).multi_processor_count // 2) * 256
            waves = max(1, round(512 * Tp / rows_per_wave))
            chunk_utterances = max(1, (waves * rows_per_wave) // Tp)
        n = min(chunk_utterances, B)
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream()
            copy = torch.cuda.Stream()
            bufs = [torch.empty(n, self.conf.in_channels, T, device=dev) for _ in range(2)]
            ready = [torch.cuda.Event() for _ in range(2)]
            free = [torch.cuda.Event() for _ in range(2)]
            copy.wait_stream(cur)
            ws_bytes = lib.vqcpc_encoder_workspace_bytes_ex(n, T, self.conf.channels, mode)
            wss = []
            for k, lo in enumerate(range(0, B, n)):
                hi, b = min(lo + n, B), k & 1
                with torch.cuda.stream(copy):
                    if k >= 2:
                        copy.wait_event(free[b])
                    bufs[b][:hi - lo].copy_(mel[lo:hi], non_blocking=True)
                    ready[b].record(copy)
                cur.wait_event(ready[b])
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)      # one per chunk: each keeps its own status word
                wss.append(ws)
                st = lib.vqcpc_encoder_forward_ex(C.byref(w), _lib.ptr(bufs[b]), hi - lo, T, _lib.ptr(ws), ws_bytes,
                                                  _lib.ptr(z[lo:hi]), None, _lib.ptr(idx[lo:hi]), None, None, mode,
                                                  _lib.current_stream_ptr())
                _lib.check(st, "Encoder.encode_from_host")
                free[b].record(cur)
            lws_bytes = lib.vqcpc_lstm_workspace_bytes(B, Tp)
            lws = torch.empty(lws_bytes, dtype=torch.uint8, device=dev)
            _lib.check(lib.vqcpc_lstm_forward_ex(C.byref(w), _lib.ptr(idx), B, Tp, _lib.ptr(lws), lws_bytes, _lib.ptr(c), mode,
                                                 _lib.current_stream_ptr()), "Encoder.encode_from_host (LSTM)")
            for ws in wss + [lws]:
                _lib.check(lib.vqcpc_check_status(_lib.ptr(ws), _lib.current_stream_ptr()), "Encoder.encode_from_host (persistent kernels)")
            for t in bufs:
                t.record_stream(copy)
        return z, c, idx

    def encode_with_aux(self, mel: Tensor):
        """``encode`` plus the pre-VQ projection (B,T',64) -- what encode.py:34-40 captures with a forward hook."""
        z, c, idx, prevq, _ = self._encode(mel, want_aux=True)
        return z, c, idx, prevq

    def _encode(self, mel: Tensor, want_aux: bool):
        _lib.require_cuda(mel, "mel")
        _check_no_grad(mel)
        if mel.dim() != 3 or mel.shape[1] != self.conf.in_channels:
            raise ValueError(f"mel must be (B, {self.conf.in_channels}, T), got {tuple(mel.shape)}")
        if mel.shape[2] < 2:
            raise ValueError("mel needs at least 2 frames (Conv1d kernel 4, stride 2, padding 1)")
        w, _keep = self.pack_weights()
        mel = mel.detach().to(torch.float32).contiguous()
        B, _, T = mel.shape
        Tp = (T - 2) // 2 + 1
        dev = mel.device
        lib = _lib.lib()
        z = torch.empty(B, Tp, self.conf.z_dim, device=dev)
        c = torch.empty(B, Tp, self.conf.c_dim, device=dev)
        idx = torch.empty(B, Tp, dtype=torch.int64, device=dev)
        prevq = torch.empty(B, Tp, self.conf.z_dim, device=dev) if want_aux else None
        hidden = torch.empty(B, Tp, self.conf.channels, device=dev) if want_aux else None
        mode = self._resolve_mode(B * Tp)
        ws_bytes = lib.vqcpc_encoder_workspace_bytes_ex(B, T, self.conf.channels, mode)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            st = lib.vqcpc_encoder_forward_ex(C.byref(w), _lib.ptr(mel), B, T, _lib.ptr(ws), ws_bytes, _lib.ptr(z),
                                              _lib.ptr(c), _lib.ptr(idx), _lib.ptr(prevq), _lib.ptr(hidden), mode,
                                              _lib.current_stream_ptr())
            _lib.check(st, "Encoder.encode")
            if B > 0:
                _lib.check(lib.vqcpc_check_status(_lib.ptr(ws), _lib.current_stream_ptr()), "Encoder.encode (persistent kernels)")
        if want_aux:
            last = self.encoder[-1]
            for hook in list(last._forward_hooks.values()):
                hook(last, (hidden,), prevq)
        return z, c, idx, prevq, hidden

    def forward(self, mels):
        raise NotImplementedError("Encoder.forward is the CPC training path (model.py:72-86): out of scope of "
                                  "the inference hot path; use .encode()")
