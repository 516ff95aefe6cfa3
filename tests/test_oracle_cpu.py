"""CPU: the oracle restatement against the reference-generated golden vectors (and, when the
reference tree is present, against the live reference itself)."""
import os

import numpy as np
import pytest
import torch

from oracle import encoder as oenc
from oracle import fixtures, mulaw, reference_loader
from oracle import vocoder as ovoc

ENCODER_CASES = ["encoder_c768_T200_init", "encoder_c512_T201_trained", "encoder_c768_T301_B3_trained",
                 "encoder_c768_T8_B2_trained"]


def load_case(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    sd = fixtures.encoder_init_state(int(g["channels"]), seed=int(g["weight_seed"]))
    if bool(g["perturbed"]):
        sd = fixtures.perturb_encoder_state(sd)
    assert abs(fixtures.state_checksum(sd) - float(g["weight_checksum"])) < 1e-6 * float(g["weight_checksum"])
    mel = fixtures.synthetic_mel(int(g["B"]), int(g["T"]), seed=int(g["mel_seed"]), kind=str(g["mel_kind"]))
    return g, sd, mel


@pytest.mark.parametrize("name", ENCODER_CASES)
def test_encoder_oracle_matches_golden(golden_dir, name):
    g, sd, mel = load_case(golden_dir, name)
    z, c, idx, z_pre = oenc.encode(sd, mel, return_aux=True)
    ref_pre = torch.from_numpy(g["z_pre"])
    assert torch.allclose(z_pre, ref_pre, rtol=1e-4, atol=2e-6), float((z_pre - ref_pre).abs().max())
    ref_idx = torch.from_numpy(g["indices"])
    # indices: identical except documented near-ties
    rep = oenc.classify_index_mismatches(ref_pre, sd["codebook.embedding"], idx, ref_idx, slack=1e-6)
    assert rep["hard"] == 0, rep
    same = (idx == ref_idx)
    assert same.float().mean() > 0.995
    ref_z = torch.from_numpy(g["z"])
    assert torch.equal(z[same], ref_z[same])           # gather of identical rows is bit-exact
    if bool(same.all()):
        ref_c = torch.from_numpy(g["c"])
        assert torch.allclose(c, ref_c, rtol=1e-4, atol=1e-5), float((c - ref_c).abs().max())


def test_vq_oracle_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "vq_lookup.npz"))
    for kind in ("init", "trained"):
        x, cb = fixtures.vq_inputs(20000, kind=kind, seed=1234, batch=2)
        q, idx = oenc.vq_lookup(x, cb)
        ref = torch.from_numpy(g[f"{kind}_indices"].astype(np.int64))
        rep = oenc.classify_index_mismatches(x, cb, idx, ref)
        assert rep["hard"] == 0, rep
        assert rep["mismatches"] <= 20
        if rep["mismatches"] == 0:
            assert abs(float(q.double().sum()) - float(g[f"{kind}_qsum"])) < 1e-9


def test_vq_exact_tie_lowest_index(golden_dir):
    g = np.load(os.path.join(golden_dir, "vq_lookup.npz"))
    x, cb = fixtures.vq_inputs(16, kind="trained", seed=5)
    cb[300] = cb[7]
    xq = cb[[7, 300, 12, 7]][None]
    _, idx = oenc.vq_lookup(xq, cb)
    assert idx.tolist() == g["tie_indices"].tolist() == [[7, 7, 12, 7]]


def test_vq_fp64_truth_agrees_on_trained_like():
    x, cb = fixtures.vq_inputs(5000, kind="trained", seed=3)
    _, idx = oenc.vq_lookup(x, cb)
    truth = oenc.vq_scores_exact(x, cb).argmin(dim=-1)
    assert torch.equal(idx.flatten(), truth)


def test_mulaw_lut_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "mulaw_lut.npz"))
    lut = mulaw.mulaw_decode_lut(8)
    assert lut.shape == (256,)
    np.testing.assert_allclose(lut, g["decode_lut"].astype(np.float32), rtol=0, atol=0)
    assert abs(lut[128] - 8.62116e-5) < 1e-9 and abs(lut[127] + 8.62116e-5) < 1e-9
    np.testing.assert_array_equal(mulaw.mulaw_encode_float(g["encode_grid"]), g["encode_codes"])
    # decode(encode(x)) stays within one quantisation cell
    codes = mulaw.mulaw_encode_float(np.linspace(-1, 1, 513)).astype(np.int64)
    assert codes.min() == 0 and codes.max() == 255


@pytest.mark.skipif(not reference_loader.reference_available(), reason="reference tree not present")
def test_init_matches_reference():
    for C in (512, 768):
        enc = reference_loader.build_reference_encoder(C, seed=13)
        ref = enc.state_dict()
        ours = fixtures.encoder_init_state(C, seed=13)
        assert set(ref.keys()) == set(ours.keys())
        for k in ref:
            assert torch.equal(ref[k], ours[k]), k


@pytest.mark.skipif(not reference_loader.reference_available(), reason="reference tree not present")
def test_oracle_matches_live_reference():
    enc = reference_loader.build_reference_encoder(512, seed=13)
    sd = fixtures.perturb_encoder_state({k: v.clone() for k, v in enc.state_dict().items()}, seed=4)
    enc.load_state_dict(sd)
    mel = fixtures.synthetic_mel(2, 57, seed=11)
    with torch.no_grad():
        z, c, idx = enc.encode(mel)
    z2, c2, idx2 = oenc.encode(sd, mel)
    assert torch.equal(idx, idx2)
    assert torch.equal(z, z2)
    assert torch.allclose(c, c2, rtol=1e-4, atol=1e-5)


# ---------------------------------------------------------------- vocoder oracle (self-consistency)
def small_vocoder():
    return ovoc.init_state_dict(n_speakers=102, seed=13)


def test_vocoder_oracle_shapes_and_embed():
    sd = small_vocoder()
    codes, spk, u = fixtures.vocoder_inputs(2, 3, seed=0, n_steps=40)
    emb = ovoc.embed_inputs(sd, codes, spk)
    assert emb.shape == (2, 6, 128)
    # x2 nearest: frames 2k and 2k+1 carry code k; speaker half is constant over time
    assert torch.equal(emb[:, 0, :64], emb[:, 1, :64])
    assert torch.equal(emb[:, 0, 64:], emb[:, 5, 64:])
    wav, x, logits = ovoc.generate(sd, codes, spk, u, n_steps=40, return_all=True)
    assert wav.shape == (2, 40) and x.shape == (2, 40) and logits.shape == (2, 40, 256)
    assert float(wav.abs().max()) <= 1.0
    # teacher forcing on the generated series reproduces the free-running logits
    x_in = torch.cat([torch.full((2, 1), 128, dtype=torch.int64), x[:, :-1]], dim=1)
    tf = ovoc.forward_teacher_forced(sd, x_in, codes, spk)
    assert torch.allclose(tf, logits, rtol=0, atol=1e-6)


def test_vocoder_oracle_matches_torch_modules():
    """The restated gate equations against torch's own nn.GRU / nn.GRUCell (the modules rnnms is built
    from): conditioning network and one AR step."""
    import torch.nn as nn
    sd = small_vocoder()
    codes, spk, _ = fixtures.vocoder_inputs(2, 5, seed=1)
    u_in = ovoc.embed_inputs(sd, codes, spk)
    gru = nn.GRU(128, 128, num_layers=2, batch_first=True, bidirectional=True)
    gru.load_state_dict({k[len("rnnms.prenet.net."):]: v for k, v in sd.items() if k.startswith("rnnms.prenet.net.")})
    with torch.no_grad():
        ref, _ = gru(u_in)
    assert torch.allclose(ovoc.prenet(sd, u_in), ref, rtol=1e-5, atol=1e-6)
    cell = nn.GRUCell(512, 896)
    cell.load_state_dict({"weight_ih": sd["rnnms.ar.rnn.weight_ih_l0"], "weight_hh": sd["rnnms.ar.rnn.weight_hh_l0"],
                          "bias_ih": sd["rnnms.ar.rnn.bias_ih_l0"], "bias_hh": sd["rnnms.ar.rnn.bias_hh_l0"]})
    h = torch.randn(2, 896) * 0.3
    x_prev = torch.tensor([5, 200])
    with torch.no_grad():
        inp = torch.cat((sd["rnnms.ar.embedding.weight"][x_prev], ref[:, 0]), dim=-1)
        h_ref = cell(inp, h)
    _, h_new = ovoc.ar_logits_step(sd, x_prev, ref[:, 0], h)
    assert torch.allclose(h_new, h_ref, rtol=1e-5, atol=1e-6)


def test_sampling_inverse_cdf_matches_distribution():
    g = torch.Generator().manual_seed(0)
    o = torch.randn(1, 256, generator=g) * 2
    u = torch.rand(20000, generator=g)
    x = ovoc.sample_inverse_cdf(o.expand(20000, -1), u)
    p = torch.softmax(o[0].double(), dim=-1)
    emp = torch.bincount(x, minlength=256).double() / 20000
    assert float((emp - p).abs().max()) < 0.01
    cdf = ovoc.cdf_bounds(o)[0]
    lo = torch.cat([torch.zeros(1, dtype=torch.float64), cdf[:-1]])
    assert bool(((u.double() >= lo[x] - 1e-6) & (u.double() <= cdf[x] + 1e-6)).all())
    # edge uniforms
    assert int(ovoc.sample_inverse_cdf(o, torch.tensor([0.0]))) == 0
    assert int(ovoc.sample_inverse_cdf(o, torch.tensor([0.99999994]))) >= 250
