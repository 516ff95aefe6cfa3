"""Checkpoint ingestion: the three container layouts of the reference end up as state_dicts the
Encoder / Vocoder accept.  CPU only: the modules are parameter containers here."""
import torch

from oracle import fixtures, vocoder as ovoc
from vectorquantizedcpc_b200 import Encoder, Vocoder, checkpoint


def _enc_sd(channels=512):
    return fixtures.encoder_init_state(channels, seed=3)


def test_cpc_training_checkpoint_layout(tmp_path):
    # train_cpc.py:23-29 writes {"encoder", "cpc", "optimizer", "scheduler", "epoch"}; convert.py:39 reads ["encoder"]
    sd = _enc_sd(512)
    path = tmp_path / "model.ckpt-7.pt"
    torch.save({"encoder": sd, "cpc": {}, "optimizer": {}, "scheduler": {}, "epoch": 7}, path)
    got = checkpoint.extract_state_dicts(path)
    assert set(got) == {"encoder"} and set(got["encoder"]) == set(sd)
    enc = checkpoint.load_encoder(path)
    assert isinstance(enc, Encoder) and not enc.training
    assert (enc.conf.channels, enc.conf.in_channels, enc.conf.n_embeddings, enc.conf.z_dim, enc.conf.c_dim) == (512, 80, 512, 64, 256)
    for k, v in enc.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_release_vocoder_layout_and_conf_inference():
    sd = ovoc.init_state_dict(seed=5)
    voc = checkpoint.load_vocoder({"vocoder": sd, "epoch": 3})
    assert isinstance(voc, Vocoder) and not voc.training
    assert voc.conf.n_speakers == sd["speaker_embedding.weight"].shape[0]
    assert voc.conf.size_i_codebook == 512 and voc.conf.dim_i_embedding + voc.conf.dim_speaker_embedding == 128
    for k, v in voc.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_lightning_vocoder_model_layout():
    # vocoder.py:41-51: VocoderModel.model = Vocoder, VocoderModel.encoder = Encoder -> "model." / "encoder." prefixes
    vsd, esd = ovoc.init_state_dict(seed=6), _enc_sd(768)
    lightning = {"state_dict": {**{"model." + k: v for k, v in vsd.items()}, **{"encoder." + k: v for k, v in esd.items()}},
                 "epoch": 1, "global_step": 10, "hyper_parameters": {}}
    got = checkpoint.extract_state_dicts(lightning)
    assert set(got) == {"encoder", "vocoder"}
    assert set(got["vocoder"]) == set(vsd) and set(got["encoder"]) == set(esd)
    enc, voc = checkpoint.load_encoder(lightning), checkpoint.load_vocoder(lightning)
    assert enc.conf.channels == 768
    assert torch.equal(voc.state_dict()["rnnms.ar.fc2.weight"], vsd["rnnms.ar.fc2.weight"])


def test_bare_state_dicts_and_errors():
    assert set(checkpoint.extract_state_dicts(_enc_sd(512))) == {"encoder"}
    assert set(checkpoint.extract_state_dicts(ovoc.init_state_dict(seed=1))) == {"vocoder"}
    try:
        checkpoint.extract_state_dicts({"optimizer": {}, "epoch": 2})
    except KeyError as e:
        assert "epoch" in str(e)
    else:
        raise AssertionError("expected KeyError")


class _HParams:            # a non-tensor object a Lightning checkpoint may pickle next to the weights
    lr = 1e-3


def test_files_load_with_the_tensor_only_unpickler_by_default(tmp_path):
    """Downloaded release checkpoints are untrusted files: the default load must not run pickled code; a checkpoint
    that really carries arbitrary objects needs the explicit trust_pickle=True."""
    import pytest
    vsd = ovoc.init_state_dict(seed=2)
    safe = tmp_path / "vocoder.pt"
    torch.save({"vocoder": vsd, "epoch": 1}, safe)
    voc = checkpoint.load_vocoder(safe)
    assert torch.equal(voc.state_dict()["rnnms.ar.fc1.bias"], vsd["rnnms.ar.fc1.bias"])
    unsafe = tmp_path / "lightning.ckpt"
    torch.save({"state_dict": {"model." + k: v for k, v in vsd.items()}, "hyper_parameters": _HParams()}, unsafe)
    with pytest.raises(RuntimeError, match="trust_pickle"):
        checkpoint.load_vocoder(unsafe)
    voc2 = checkpoint.load_vocoder(unsafe, trust_pickle=True)
    assert torch.equal(voc2.state_dict()["rnnms.ar.fc1.bias"], vsd["rnnms.ar.fc1.bias"])
