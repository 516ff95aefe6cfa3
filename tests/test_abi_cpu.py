"""CPU: the C-ABI library loads and exports every symbol include/vqcpc.h declares; host-side argument
checking and the no-fallback rule (no compute calls here -- there is no GPU)."""
import os
import re

import pytest
import torch

from vectorquantizedcpc_b200 import ConfEncoder, ConfVocoder, Encoder, Vocoder, VQEmbeddingEMA, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vqcpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vqcpc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        from vectorquantizedcpc_b200 import build
        build.build()
    lib = _lib.lib()
    declared = header_symbols()
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vqcpc.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared
    assert lib.vqcpc_abi_version() == 1


def test_struct_layout_matches_header(tmp_path):
    """The ctypes mirrors against the header itself: gcc compiles include/vqcpc.h and prints sizeof / offsetof."""
    import ctypes as C
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "vqcpc.h"\n'
        'int main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(vqcpc_encoder_weights), sizeof(vqcpc_vocoder_weights),\n'
        '  offsetof(vqcpc_encoder_weights, lstm_whh_p), offsetof(vqcpc_encoder_weights, lstm_table),\n'
        '  offsetof(vqcpc_vocoder_weights, mulaw_lut), sizeof(vqcpc_logmel_config)); return 0; }\n')
    exe = tmp_path / "layout"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)])
    enc_sz, voc_sz, off_whh, off_tab, off_lut, mel_sz = map(int, subprocess.check_output([str(exe)]).split())
    assert C.sizeof(_lib.EncoderWeights) == enc_sz == 24 + 8 * 29      # 6 int32 + 29 pointers
    assert C.sizeof(_lib.VocoderWeights) == voc_sz == 24 + 8 * 21
    assert _lib.EncoderWeights.lstm_whh_p.offset == off_whh and _lib.EncoderWeights.lstm_table.offset == off_tab
    assert _lib.VocoderWeights.mulaw_lut.offset == off_lut
    assert C.sizeof(_lib.LogMelConfig) == mel_sz


def test_state_dict_layout_matches_reference_appendix_c():
    enc = Encoder(ConfEncoder(channels=768))
    sd = enc.state_dict()
    assert sd["conv.weight"].shape == (768, 80, 4)
    for i in (0, 3, 6, 9, 12):
        assert sd[f"encoder.{i}.weight"].shape == (768,) and sd[f"encoder.{i}.bias"].shape == (768,)
    for i in (2, 5, 8, 11):
        assert sd[f"encoder.{i}.weight"].shape == (768, 768)
    assert sd["encoder.14.weight"].shape == (64, 768) and sd["encoder.14.bias"].shape == (64,)
    assert sd["codebook.embedding"].shape == (512, 64) and sd["codebook.ema_count"].shape == (512,)
    assert sd["rnn.weight_ih_l0"].shape == (1024, 64) and sd["rnn.weight_hh_l0"].shape == (1024, 256)
    assert sum(p.numel() for p in enc.parameters()) == 2991680
    enc512 = Encoder(channels=512, in_channels=80, n_embeddings=512, z_dim=64, c_dim=256)   # kwargs style
    assert sum(p.numel() for p in enc512.parameters()) == 1580096
    voc = Vocoder(ConfVocoder())
    vsd = voc.state_dict()
    assert vsd["code_embedding.weight"].shape == (512, 64) and vsd["speaker_embedding.weight"].shape == (102, 64)
    assert vsd["rnnms.ar.rnn.weight_hh_l0"].shape == (2688, 896) and vsd["rnnms.ar.rnn.weight_ih_l0"].shape == (2688, 512)
    assert vsd["rnnms.prenet.net.weight_ih_l1_reverse"].shape == (384, 256)
    ar = sum(v.numel() for k, v in vsd.items() if k.startswith("rnnms.ar."))
    assert ar == 4151040


def test_vocoder_checkpoint_key_remap():
    voc = Vocoder()
    sd = voc.state_dict()
    alt = {}
    for k, v in sd.items():
        k2 = k.replace("rnnms.prenet.net.", "rnnms.rnn1.").replace("rnnms.ar.rnn.", "rnnms.rnn2.") \
              .replace("rnnms.ar.embedding.", "rnnms.embedding.").replace("rnnms.ar.fc", "rnnms.fc")
        alt["model." + k2] = v.clone()
    voc2 = Vocoder()
    voc2.load_state_dict(alt)
    for k, v in voc2.state_dict().items():
        assert torch.equal(v, sd[k])


def test_no_cpu_fallback():
    enc = Encoder(ConfEncoder(channels=512))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc.encode(torch.zeros(1, 80, 16))
    vq = VQEmbeddingEMA(512, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vq.encode(torch.zeros(1, 4, 64))
    voc = Vocoder()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        voc.generate(torch.zeros(1, 2, dtype=torch.int64), torch.zeros(1, dtype=torch.int64))
    with pytest.raises(NotImplementedError):
        enc.forward(torch.zeros(1, 80, 16))


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "vectorquantizedcpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_unsupported_vocoder_dims_rejected():
    from vectorquantizedcpc_b200.network_vocoder import ConfRNNMSVocoder, ConfWaveAR
    with pytest.raises(ValueError):
        Vocoder(ConfVocoder(rnnms=ConfRNNMSVocoder(wave_ar=ConfWaveAR(size_h_rnn=512))))
