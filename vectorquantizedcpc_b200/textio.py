"""File outputs of ``encode.py`` / ``convert.py``.

* ``save_txt`` / ``format_txt`` -- ``np.savetxt(file, z, fmt="%.16f")`` of encode.py:48-52,57-67, formatted on the GPU
  (``csrc/textdump.cu``): byte-identical to numpy's output, one device->host copy of the finished text.
* ``write_wav`` -- ``librosa.output.write_wav(path, output.astype(np.float32), sr)`` of convert.py:83, i.e.
  ``scipy.io.wavfile.write`` of a float32 array: a RIFF/WAVE file with format tag 3 (IEEE float), an 18-byte ``fmt `` chunk
  and a ``fact`` chunk.  Host code: it is the file write itself, there is nothing to compute."""
from __future__ import annotations

import ctypes as C
import struct

import torch
from torch import Tensor

from . import _lib


def format_txt(x: Tensor) -> Tensor:
    """(rows, cols) or (rows,) fp32 CUDA tensor -> uint8 CUDA tensor holding exactly what ``np.savetxt(f, x.cpu().numpy(),
    fmt="%.16f")`` writes (a 1-D input is one value per line, as numpy does)."""
    _lib.require_cuda(x, "x")
    if x.dim() == 1:
        x = x[:, None]
    if x.dim() != 2:
        raise ValueError(f"format_txt expects a 1-D or 2-D tensor, got {tuple(x.shape)}")     # np.savetxt's own restriction
    x = x.detach().to(torch.float32).contiguous()
    rows, cols = x.shape
    if rows == 0 or cols == 0:
        return torch.empty(0, dtype=torch.uint8, device=x.device)
    lib = _lib.lib()
    with torch.cuda.device(x.device):
        ws_bytes = lib.vqcpc_textdump_workspace_bytes(rows, cols)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        n = C.c_int64(0)
        _lib.check(lib.vqcpc_textdump_f16(_lib.ptr(x), rows, cols, None, 0, C.byref(n), _lib.ptr(ws), ws_bytes,
                                          _lib.current_stream_ptr()), "format_txt (length)")
        out = torch.empty(n.value, dtype=torch.uint8, device=x.device)
        _lib.check(lib.vqcpc_textdump_f16(_lib.ptr(x), rows, cols, _lib.ptr(out), n.value, C.byref(n), _lib.ptr(ws), ws_bytes,
                                          _lib.current_stream_ptr()), "format_txt")
    return out


def save_txt(path, x: Tensor) -> int:
    """Drop-in for ``np.savetxt(path, x, fmt="%.16f")`` (encode.py:51,59,66) with ``x`` still on the GPU.  Returns bytes written."""
    data = format_txt(x).cpu().numpy().tobytes()
    with open(path, "wb") as f:
        f.write(data)
    return len(data)


def wav_bytes(wav, sr: int) -> bytes:
    """The bytes ``scipy.io.wavfile.write(path, sr, wav.astype(float32))`` produces for a mono (N,) or (N, channels) signal."""
    t = torch.as_tensor(wav).detach().to("cpu", torch.float32).contiguous()
    if t.dim() not in (1, 2):
        raise ValueError("wav must be (N,) or (N, channels)")
    channels = 1 if t.dim() == 1 else t.shape[1]
    n_frames = t.shape[0]
    data = t.numpy().astype("<f4", copy=False).tobytes()
    bits = 32
    block_align = channels * bits // 8
    fmt = struct.pack("<HHIIHH", 0x0003, channels, int(sr), int(sr) * block_align, block_align, bits) + b"\x00\x00"   # cbSize = 0
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt
    body += b"fact" + struct.pack("<II", 4, n_frames)
    body += b"data" + struct.pack("<I", len(data)) + data
    if len(data) % 2:
        body += b"\x00"
    if len(body) + 8 > 0xFFFFFFFF:
        raise ValueError("Data exceeds wave file size limit")
    return b"RIFF" + struct.pack("<I", len(body)) + body


def write_wav(path, wav, sr: int = 16000) -> None:
    """``librosa.output.write_wav(path, wav.astype(np.float32), sr=sr)`` (convert.py:83; removed from librosa 0.8): a 32-bit
    float WAVE file.  ``wav`` may live on the GPU (one device->host copy)."""
    with open(path, "wb") as f:
        f.write(wav_bytes(wav, sr))
