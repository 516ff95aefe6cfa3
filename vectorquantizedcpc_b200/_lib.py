"""ctypes binding of libvqcpc_b200.so -- the C ABI declared in include/vqcpc.h.

There is NO CPU fallback: loading fails loudly if the library has not been built
(``python -m vectorquantizedcpc_b200.build``) and every call raises on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvqcpc_b200.so")

ERR_ARG, ERR_CUDA, ERR_DEVICE, ERR_TIMEOUT, ERR_INDEX = 1, 2, 3, 4, 5
GEMM_FP32, GEMM_BF16X3, GEMM_BF16 = 0, 1, 2
f32p = C.POINTER(C.c_float)


class EncoderWeights(C.Structure):
    """``vqcpc_encoder_weights`` (include/vqcpc.h)."""
    _fields_ = [
        ("in_channels", C.c_int32), ("channels", C.c_int32), ("n_embeddings", C.c_int32),
        ("z_dim", C.c_int32), ("c_dim", C.c_int32), ("_pad", C.c_int32),
        ("conv_w", C.c_void_p),
        ("ln_w", C.c_void_p * 5), ("ln_b", C.c_void_p * 5), ("fc_w", C.c_void_p * 4),
        ("proj_w", C.c_void_p), ("proj_b", C.c_void_p), ("codebook", C.c_void_p),
        ("lstm_w_ih", C.c_void_p), ("lstm_w_hh", C.c_void_p), ("lstm_b", C.c_void_p),
        ("conv_wp", C.c_void_p), ("fc_wp", C.c_void_p * 4), ("proj_wp", C.c_void_p), ("lstm_whh_p", C.c_void_p),
        ("lstm_table", C.c_void_p),
    ]


class VocoderWeights(C.Structure):
    """``vqcpc_vocoder_weights`` (include/vqcpc.h)."""
    _fields_ = [
        ("n_codes", C.c_int32), ("dim_code", C.c_int32), ("n_speakers", C.c_int32),
        ("dim_speaker", C.c_int32), ("upsample_t", C.c_int32), ("_pad", C.c_int32),
        ("code_emb", C.c_void_p), ("spk_emb", C.c_void_p),
        ("pre_w_ih", C.c_void_p * 2), ("pre_b_ih", C.c_void_p * 2),
        ("pre_w_hh", C.c_void_p * 2), ("pre_b_hh", C.c_void_p * 2),
        ("ar_w_ih", C.c_void_p), ("ar_b_ih", C.c_void_p), ("ar_w_hh", C.c_void_p), ("ar_b_hh", C.c_void_p),
        ("fc1_w", C.c_void_p), ("fc1_b", C.c_void_p), ("fc2_w", C.c_void_p), ("fc2_b", C.c_void_p),
        ("ar_emb", C.c_void_p), ("eprime", C.c_void_p), ("mulaw_lut", C.c_void_p),
    ]


# name -> (restype, argtypes); must list EVERY symbol include/vqcpc.h declares (tests/test_abi_cpu.py checks)
class LogMelConfig(C.Structure):
    """``vqcpc_logmel_config`` (include/vqcpc.h)."""
    _fields_ = [("n_fft", C.c_int32), ("win_length", C.c_int32), ("hop_length", C.c_int32), ("n_mels", C.c_int32),
                ("n_freq_padded", C.c_int32), ("preemph", C.c_float), ("top_db", C.c_float)]


_vp, _i32, _i64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
SIGNATURES = {
    "vqcpc_last_error": (C.c_char_p, []),
    "vqcpc_abi_version": (C.c_int, []),
    "vqcpc_launch_count": (C.c_uint64, []),
    "vqcpc_device_check": (C.c_int, [C.c_int]),
    "vqcpc_linear_f32": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i64, _i32, _i32, _vp]),
    "vqcpc_linear_tc": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "vqcpc_split_planes": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp]),
    "vqcpc_layernorm_relu_f32": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp]),
    "vqcpc_vq_lookup": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "vqcpc_vq_workspace_bytes": (_sz, []),
    "vqcpc_encoder_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "vqcpc_encoder_forward": (C.c_int, [C.POINTER(EncoderWeights), _vp, _i32, _i32, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vqcpc_encoder_workspace_bytes_ex": (_sz, [_i32, _i32, _i32, _i32]),
    "vqcpc_encoder_forward_ex": (C.c_int, [C.POINTER(EncoderWeights), _vp, _i32, _i32, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _i32, _vp]),
    "vqcpc_lstm_workspace_bytes": (_sz, [_i32, _i32]),
    "vqcpc_lstm_forward": (C.c_int, [C.POINTER(EncoderWeights), _vp, _i32, _i32, _vp, _sz, _vp, _vp]),
    "vqcpc_lstm_forward_ex": (C.c_int, [C.POINTER(EncoderWeights), _vp, _i32, _i32, _vp, _sz, _vp, _i32, _vp]),
    "vqcpc_vocoder_pack": (C.c_int, [C.POINTER(VocoderWeights), _vp, _vp]),
    "vqcpc_vocoder_workspace_bytes": (_sz, [_i32, _i32]),
    "vqcpc_vocoder_condition": (C.c_int, [C.POINTER(VocoderWeights), _vp, _vp, _i32, _i32, _vp, _sz, _vp, _vp, _vp]),
    "vqcpc_textdump_workspace_bytes": (_sz, [C.c_int64, _i32]),
    "vqcpc_textdump_f16": (C.c_int, [_vp, C.c_int64, _i32, _vp, _sz, C.POINTER(C.c_int64), _vp, _sz, _vp]),
    "vqcpc_loudness_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "vqcpc_integrated_loudness": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _sz, _vp, _vp]),
    "vqcpc_loudness_normalize": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _sz, _vp, _vp, _vp]),
    "vqcpc_logmel_workspace_bytes": (_sz, [C.POINTER(LogMelConfig), _i32, _i32]),
    "vqcpc_logmel_forward": (C.c_int, [C.POINTER(LogMelConfig), _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "vqcpc_vocoder_condition_ragged": (C.c_int, [C.POINTER(VocoderWeights), _vp, _vp, _vp, _i32, _i32, _vp, _sz, _vp, _vp, _vp]),
    "vqcpc_vocoder_generate": (C.c_int, [C.POINTER(VocoderWeights), _vp, _vp, _i32, _i32, _i32, _vp, _sz, _vp, _vp, _vp, _vp]),
    "vqcpc_vocoder_logits_tf": (C.c_int, [C.POINTER(VocoderWeights), _vp, _vp, _i32, _i32, _i32, _vp, _sz, _vp, _vp]),
    "vqcpc_check_status": (C.c_int, [_vp, _vp]),
    "vqcpc_debug_set_ar_trace": (C.c_int, [_vp, _i32, _i32, _i32]),
    "vqcpc_debug_set_ar_poll_gap": (C.c_int, [_i32]),
    "vqcpc_debug_set_ar_cluster": (C.c_int, [_i32, _i32, _i32]),
    "vqcpc_ar_cluster_active": (C.c_int, []),
    "vqcpc_debug_exchange_floor": (C.c_int, [_vp, _sz, _i32, C.POINTER(C.c_double), _vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise if it is missing -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m vectorquantizedcpc_b200.build` "
            "(nvcc, sm_100a).  vectorquantizedcpc_b200 has no CPU / PyTorch fallback path.")
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


class VqcpcError(RuntimeError):
    pass


def check(status: int, what: str) -> None:
    """Raise on a non-zero C-ABI status: ValueError for argument errors, RuntimeError otherwise."""
    if status == 0:
        return
    msg = lib().vqcpc_last_error().decode("utf-8", "replace")
    if status == ERR_ARG:
        raise ValueError(f"{what}: {msg}")
    if status == ERR_INDEX:
        raise IndexError(f"{what}: {msg} (nn.Embedding would raise)")
    raise VqcpcError(f"{what}: status {status}: {msg}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t, name: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: vectorquantizedcpc_b200 runs only on a CUDA (sm_100a) device -- "
            "there is no CPU fallback; move the module and its inputs with .to('cuda').")
