"""GPU parity tests: the sm_100a CUDA path (through the C ABI / the reference-shaped Python classes) against
the CPU oracle and the reference-generated golden vectors.  Run on the B200 box: pytest -m gpu."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import encoder as oenc
from oracle import fixtures, mulaw
from oracle import vocoder as ovoc
from vectorquantizedcpc_b200 import ConfEncoder, Encoder, Vocoder, VQEmbeddingEMA, _lib

pytestmark = pytest.mark.gpu

# Tolerances (fp32 parity path).  north_star: encoder z/c within 1e-4 relative in fp32.
RTOL = 1e-4
ATOL_ZPRE = 2e-5
ATOL_C = 2e-5
ATOL_LOGITS = 2e-4     # teacher-forced logits, absolute (logits are O(0.1 .. 1) at random init)


def dev():
    return torch.device("cuda:0")


def make_encoder(C_, perturbed, seed=13):
    sd = fixtures.encoder_init_state(C_, seed=seed)
    if perturbed:
        sd = fixtures.perturb_encoder_state(sd)
    enc = Encoder(ConfEncoder(channels=C_))
    enc.load_state_dict(sd)
    return enc.to(dev()).eval(), sd


_VOC = {}


def make_vocoder():
    if "v" not in _VOC:
        sd = ovoc.init_state_dict(seed=13)
        v = Vocoder()
        v.load_state_dict(sd)
        _VOC["v"] = (v.to(dev()).eval(), sd)
    return _VOC["v"]


# ------------------------------------------------------------------------------------------ building blocks
@pytest.mark.parametrize("M,N,K,bias", [(100, 768, 320, False), (1, 64, 768, True), (1000, 2688, 256, True),
                                        (40000, 768, 768, False), (257, 1024, 64, True), (0, 64, 64, False)])
def test_linear_f32(M, N, K, bias):
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) if bias else None
    ref = (A.double() @ W.double().t() + (b.double() if bias else 0)).float()
    Ad, Wd = A.to(dev()), W.to(dev())
    bd = b.to(dev()) if bias else None
    out = torch.empty(M, N, device=dev())
    st = _lib.lib().vqcpc_linear_f32(_lib.ptr(Ad), K, _lib.ptr(Wd), K, _lib.ptr(bd), _lib.ptr(out), N, M, N, K,
                                     _lib.current_stream_ptr())
    _lib.check(st, "linear")
    torch.cuda.synchronize()
    assert torch.allclose(out.cpu(), ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("M,N,K,bias", [(1000, 768, 320, False), (300, 64, 768, True), (128, 512, 512, False),
                                        (1, 256, 64, True), (40000, 768, 768, False), (4097, 128, 128, True)])
def test_linear_tensor_core_split(M, N, K, bias):
    """tcgen05 path with the bf16 hi/lo split (3 MMA terms): fp32-grade accuracy against an fp64 product."""
    g = torch.Generator().manual_seed(M + N + K + 1)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g) if bias else None
    ref = (A.double() @ W.double().t() + (b.double() if bias else 0)).float()
    Ad, Wd = A.to(dev()), W.to(dev())
    bd = b.to(dev()) if bias else None
    out = torch.full((M, N), float("nan"), device=dev())
    ap = torch.empty(M, 2 * K, dtype=torch.bfloat16, device=dev())
    wp = torch.empty(N, 2 * K, dtype=torch.bfloat16, device=dev())
    err = torch.zeros(1, dtype=torch.int32, device=dev())
    st = _lib.lib().vqcpc_linear_tc(_lib.ptr(Ad), _lib.ptr(Wd), _lib.ptr(bd), _lib.ptr(out), M, N, K, 3, _lib.ptr(ap),
                                    _lib.ptr(wp), _lib.ptr(err), _lib.current_stream_ptr())
    _lib.check(st, "linear_tc")
    torch.cuda.synchronize()
    assert int(err) == 0, "tensor-core pipeline timed out"
    e = float((out.cpu() - ref).abs().max())
    print(f"[linear_tc {M}x{N}x{K}] max abs err {e:.2e}")
    assert torch.allclose(out.cpu(), ref, rtol=1e-4, atol=3e-5), e


def test_linear_rejects_bad_k():
    A = torch.zeros(4, 20, device=dev())
    out = torch.empty(4, 8, device=dev())
    st = _lib.lib().vqcpc_linear_f32(_lib.ptr(A), 20, _lib.ptr(A), 20, None, _lib.ptr(out), 8, 4, 8, 20,
                                     _lib.current_stream_ptr())
    with pytest.raises(ValueError):
        _lib.check(st, "linear")


@pytest.mark.parametrize("rows,C_", [(100, 768), (7, 512), (4097, 768), (3, 128)])
def test_layernorm_relu(rows, C_):
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, C_, generator=g) * 3 + 0.5
    w = torch.randn(C_, generator=g)
    b = torch.randn(C_, generator=g)
    ref = torch.relu(oenc.layer_norm(x.double(), w.double(), b.double())).float()
    xd, wd, bd = x.to(dev()), w.to(dev()), b.to(dev())
    _lib.check(_lib.lib().vqcpc_layernorm_relu_f32(_lib.ptr(xd), _lib.ptr(wd), _lib.ptr(bd), rows, C_,
                                                   _lib.current_stream_ptr()), "ln")
    assert torch.allclose(xd.cpu(), ref, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------ VQ lookup
def run_vq(x, cb):
    vq = VQEmbeddingEMA(512, 64)
    vq.embedding.copy_(cb)
    vq = vq.to(dev())
    q, idx = vq.encode(x.to(dev()))
    return q.cpu(), idx.cpu()


@pytest.mark.parametrize("kind", ["init", "trained"])
def test_vq_lookup_matches_oracle_and_golden(golden_dir, kind):
    g = np.load(os.path.join(golden_dir, "vq_lookup.npz"))
    x, cb = fixtures.vq_inputs(20000, kind=kind, seed=1234, batch=2)
    q, idx = run_vq(x, cb)
    assert idx.dtype == torch.int64 and idx.shape == (2, 20000) and q.shape == x.shape
    _, idx_o = oenc.vq_lookup(x, cb)
    rep = oenc.classify_index_mismatches(x, cb, idx, idx_o)
    assert rep["hard"] == 0, rep                     # zero non-near-tie mismatches vs the oracle
    rep_g = oenc.classify_index_mismatches(x, cb, idx, torch.from_numpy(g[f"{kind}_indices"].astype(np.int64)))
    assert rep_g["hard"] == 0, rep_g                 # ... and vs the live reference's golden indices
    assert rep["mismatches"] <= 20 and rep_g["mismatches"] <= 20
    # against fp64 truth the kernel (which drops the argmin-invariant |x|^2 term) must be at least as good
    truth = oenc.vq_scores_exact(x, cb).argmin(dim=-1)
    rep_t = oenc.classify_index_mismatches(x, cb, idx, truth.view(2, -1))
    assert rep_t["hard"] == 0 and rep_t["mismatches"] <= rep_g["mismatches"] + 2, (rep_t, rep_g)
    assert torch.equal(q, cb[idx])                   # gather is bit-exact
    print(f"[vq {kind}] vs oracle {rep} | vs golden {rep_g} | vs fp64 {rep_t}")


@pytest.mark.parametrize("kind", ["init", "trained"])
def test_vq_tensor_core_path_equals_fp32_path(kind):
    """n >= 8192 runs the tcgen05 coarse pass + exact recheck; it must reproduce the fp32 kernel's indices exactly
    (compared by running the same frames in chunks below the threshold), including a ragged tail and exact ties."""
    n = 50_000 + 37
    x, cb = fixtures.vq_inputs(n, kind=kind, seed=77)
    cb[300] = cb[7]                                   # exact duplicate code: ties must go to index 7
    x[0, :5] = cb[7]
    vq = VQEmbeddingEMA(512, 64)
    vq.embedding.copy_(cb)
    vq = vq.to(dev())
    xd = x.to(dev())
    q, idx = vq.encode(xd)
    parts = [vq.encode(xd[:, i:i + 4096])[1] for i in range(0, n, 4096)]
    idx_ref = torch.cat(parts, dim=1)
    nbad = int((idx != idx_ref).sum())
    print(f"[vq tc vs fp32, {kind}] mismatches {nbad} / {n}")
    assert nbad == 0
    assert idx[0, :5].tolist() == [7] * 5
    assert torch.equal(q, cb.to(dev())[idx])


def test_vq_exact_tie_and_ragged_sizes():
    x, cb = fixtures.vq_inputs(16, kind="trained", seed=5)
    cb[300] = cb[7]
    xq = cb[[7, 300, 12, 7]][None]
    _, idx = run_vq(xq, cb)
    assert idx.tolist() == [[7, 7, 12, 7]]           # lowest index wins (torch.argmin semantics)
    for n in (1, 127, 128, 129, 1000):
        x, cb = fixtures.vq_inputs(n, kind="trained", seed=n)
        q, idx = run_vq(x, cb)
        _, idx_o = oenc.vq_lookup(x, cb)
        assert torch.equal(idx, idx_o) and torch.equal(q, cb[idx])
    q, idx = run_vq(torch.zeros(2, 0, 64), cb)
    assert q.shape == (2, 0, 64) and idx.shape == (2, 0)


def test_vq_full_size_properties():
    """BASELINE config 2 size (1 M frames): idempotence, index range, checksum of checksums."""
    x, cb = fixtures.vq_inputs(1_000_000, kind="trained", seed=1234)
    vq = VQEmbeddingEMA(512, 64)
    vq.embedding.copy_(cb)
    vq = vq.to(dev())
    xd = x.to(dev())
    q, idx = vq.encode(xd)
    assert int(idx.min()) >= 0 and int(idx.max()) < 512
    q2, idx2 = vq.encode(q)                          # quantising a code vector returns the same code
    cbd = cb.to(dev())
    assert torch.equal(q2, q)
    d_self = ((cbd[idx2] - cbd[idx]) ** 2).sum(-1)
    assert float(d_self.max()) == 0.0
    assert torch.equal(q, cbd[idx])
    # every frame's chosen code is at least as close as 64 random other codes (fp64 check on a sample)
    sel = torch.randint(0, 1_000_000, (2000,))
    sc = oenc.vq_scores_exact(x[:, sel], cb)
    best = sc.min(dim=-1).values
    chosen = sc[torch.arange(2000), idx.cpu()[0, sel]]
    assert float((chosen - best).max()) < 1e-4


# ------------------------------------------------------------------------------------------ encoder
CASES = ["encoder_c768_T200_init", "encoder_c512_T201_trained", "encoder_c768_T301_B3_trained",
         "encoder_c768_T8_B2_trained"]


@pytest.mark.parametrize("name", CASES)
def test_encoder_matches_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    enc, sd = make_encoder(int(g["channels"]), bool(g["perturbed"]))
    mel = fixtures.synthetic_mel(int(g["B"]), int(g["T"]), seed=int(g["mel_seed"]), kind=str(g["mel_kind"]))
    z, c, idx, prevq = [t.cpu() for t in enc.encode_with_aux(mel.to(dev()))]
    ref_pre = torch.from_numpy(g["z_pre"])
    assert z.shape == ref_pre.shape and idx.dtype == torch.int64
    assert torch.allclose(prevq, ref_pre, rtol=RTOL, atol=ATOL_ZPRE), float((prevq - ref_pre).abs().max())
    ref_idx = torch.from_numpy(g["indices"])
    # near-tie slack: |dz| * |e_a - e_b| bound from the pre-VQ tolerance
    rep = oenc.classify_index_mismatches(ref_pre, sd["codebook.embedding"], idx, ref_idx, slack=1e-5)
    assert rep["hard"] == 0, rep
    same = idx == ref_idx
    assert same.float().mean() >= 0.98, rep
    assert torch.equal(z[same], torch.from_numpy(g["z"])[same])
    ref_c = torch.from_numpy(g["c"])
    ok_utts = same.all(dim=1)
    assert torch.allclose(c[ok_utts], ref_c[ok_utts], rtol=RTOL, atol=ATOL_C), float((c[ok_utts] - ref_c[ok_utts]).abs().max())
    print(f"[{name}] idx mismatches {rep}; max|dz_pre| {float((prevq - ref_pre).abs().max()):.2e}; "
          f"max|dc| {float((c[ok_utts] - ref_c[ok_utts]).abs().max()) if ok_utts.any() else float('nan'):.2e}")


@pytest.mark.parametrize("C_,B,T", [(768, 8, 300), (512, 33, 101), (512, 70, 61)])
def test_encoder_tensor_core_mode_matches_oracle(C_, B, T):
    """tcgen05 GEMMs with the bf16 hi/lo split: pre-VQ z within 1e-4 relative (the north-star fp32 bound); indices
    equal to the oracle's except near-ties of the size the z error allows; c compared where the indices agree."""
    enc, sd = make_encoder(C_, True)
    enc.gemm_mode = "bf16x3"
    mel = fixtures.synthetic_mel(B, T, seed=21)
    z, c, idx, prevq = [t.cpu() for t in enc.encode_with_aux(mel.to(dev()))]
    enc.gemm_mode = "fp32"
    z32, c32, idx32, prevq32 = [t.cpu() for t in enc.encode_with_aux(mel.to(dev()))]
    zo, co, io, zp = oenc.encode(sd, mel, return_aux=True)
    scale = float(zp.abs().max())
    err = float((prevq - zp).abs().max())
    print(f"[encoder bf16x3 C={C_}] max|dz_pre| = {err:.2e} (scale {scale:.2f}; fp32 path {float((prevq32 - zp).abs().max()):.2e})")
    assert err <= 1e-4 * scale + 1e-5
    rep = oenc.classify_index_mismatches(zp, sd["codebook.embedding"], idx, io, slack=4 * err * 2.0)
    assert rep["hard"] == 0, rep
    same = idx == io
    assert same.float().mean() > 0.99
    assert torch.equal(z[same], zo[same])
    ok = same.all(dim=1)
    assert ok.any()
    assert torch.allclose(c[ok], co[ok], rtol=RTOL, atol=ATOL_C)
    assert torch.equal(idx32, io)


@pytest.mark.parametrize("C_,B,T", [(768, 40, 300), (512, 33, 101)])
def test_encoder_bf16_speed_mode_stated_bound(C_, B, T):
    """gemm_mode = "bf16" (north_star: "bf16 speed mode with a stated looser bound"; SURVEY 7.2 "ship two modes"): the conv /
    MLP / projection products run as single-pass bf16 tcgen05 MMAs (fp32 accumulate, fp32 LayerNorm statistics).  Stated bound,
    asserted here: pre-VQ z within BF16_Z_REL of max|z| of the oracle; every index either equals the oracle's or is a near-tie
    of the size that z error allows (no hard mismatch); >= BF16_IDX_AGREE of the indices equal; z is EXACTLY the codebook row
    of the returned index; and c -- the LSTM stays in the fp32-grade bf16x3 arithmetic -- is within the fp32 tolerance for
    every utterance whose indices all agree (c is a function of the indices only)."""
    BF16_Z_REL, BF16_IDX_AGREE = 3e-2, 0.97
    enc, sd = make_encoder(C_, True)
    enc.gemm_mode = "bf16"
    mel = fixtures.synthetic_mel(B, T, seed=23)
    z, c, idx, prevq = [t.cpu() for t in enc.encode_with_aux(mel.to(dev()))]
    zo, co, io, zp = oenc.encode(sd, mel, return_aux=True)
    scale = float(zp.abs().max())
    err = float((prevq - zp).abs().max())
    agree = float((idx == io).float().mean())
    print(f"[encoder bf16 C={C_}] max|dz_pre| = {err:.3e} (scale {scale:.2f}, rel {err / scale:.2e}); index agreement {agree:.4f}")
    assert err <= BF16_Z_REL * scale
    rep = oenc.classify_index_mismatches(zp, sd["codebook.embedding"], idx, io, slack=4 * err * 2.0)
    assert rep["hard"] == 0, rep
    assert agree >= BF16_IDX_AGREE
    assert torch.equal(z, sd["codebook.embedding"][idx])
    ok = (idx == io).all(dim=1)
    if ok.any():
        assert torch.allclose(c[ok], co[ok], rtol=RTOL, atol=ATOL_C)
    # the LSTM of the speed mode is a function of the indices only: feeding ITS indices to the oracle LSTM reproduces c
    c_from_idx = oenc.lstm(sd["codebook.embedding"][idx], sd["rnn.weight_ih_l0"], sd["rnn.weight_hh_l0"], sd["rnn.bias_ih_l0"],
                           sd["rnn.bias_hh_l0"])
    assert torch.allclose(c, c_from_idx, rtol=RTOL, atol=ATOL_C)


def test_encoder_full_size_properties():
    """BASELINE configs[3] at full size (4096 utterances x 3 s = 614 400 frames, C = 768, tensor-core mode): size-
    independent properties plus an oracle check on a sample of utterances.
      * every index is a valid code and z is exactly the gathered codebook row (gather consistency);
      * |c| < 1 (an LSTM output), no NaN;
      * batch independence: utterances re-encoded alone (fp32 parity path) give the same indices except near-ties of
        the size the bf16x3 error allows, and c within tolerance where the indices agree;
      * the sampled utterances match the CPU oracle the same way."""
    enc, sd = make_encoder(768, True)
    B, T = 4096, 300
    mel = fixtures.synthetic_mel(B, T, seed=0)
    md = mel.to(dev())
    z, c, idx, prevq = enc.encode_with_aux(md)
    Tp = (T - 2) // 2 + 1
    assert z.shape == (B, Tp, 64) and c.shape == (B, Tp, 256) and idx.shape == (B, Tp) and idx.dtype == torch.int64
    assert int(idx.min()) >= 0 and int(idx.max()) < 512
    cb = enc.codebook.embedding
    assert torch.equal(z, cb[idx])
    assert bool(torch.isfinite(c).all()) and float(c.abs().max()) < 1.0
    sel = [0, 1, 2047, 4095]
    enc.gemm_mode = "fp32"
    z1, c1, i1, p1 = enc.encode_with_aux(md[sel])
    enc.gemm_mode = "auto"
    zo, co, io, zp = oenc.encode(sd, mel[sel], return_aux=True)
    assert torch.equal(i1.cpu(), io)                                   # fp32 path == oracle, bit-exact indices
    scale = float(zp.abs().max())
    err = float((prevq[sel].cpu() - zp).abs().max())
    print(f"[encode 4096x3s] max|dz_pre| (bf16x3 vs oracle, 4 utterances) = {err:.2e}, scale {scale:.2f}")
    assert err <= 1e-4 * scale + 1e-5
    rep = oenc.classify_index_mismatches(zp, sd["codebook.embedding"], idx[sel].cpu(), io, slack=4 * err * 2.0)
    assert rep["hard"] == 0, rep
    ok = (idx[sel].cpu() == io).all(dim=1)
    assert ok.any()
    assert torch.allclose(c[sel].cpu()[ok], co[ok], rtol=RTOL, atol=ATOL_C)


def test_encoder_matches_oracle_other_shapes_and_hook():
    enc, sd = make_encoder(512, True)
    aux = []
    h = enc.encoder[-1].register_forward_hook(lambda m, i, o: aux.append((i[0].clone(), o.clone())))
    for (B, T, seed) in [(1, 2, 5), (5, 33, 6), (2, 100, 7)]:
        mel = fixtures.synthetic_mel(B, T, seed=seed)
        z, c, idx = [t.cpu() for t in enc.encode(mel.to(dev()))]
        zo, co, io, zp = oenc.encode(sd, mel, return_aux=True)
        hid_in, pre = aux.pop()
        assert hid_in.shape == (B, (T - 2) // 2 + 1, 512)
        assert torch.allclose(pre.cpu(), zp, rtol=RTOL, atol=ATOL_ZPRE)
        assert torch.equal(idx, io) and torch.equal(z, zo)
        assert torch.allclose(c, co, rtol=RTOL, atol=ATOL_C)
    h.remove()
    with pytest.raises(ValueError):
        enc.encode(torch.zeros(1, 40, 10, device=dev()))
    with pytest.raises(ValueError):
        enc.encode(torch.zeros(1, 80, 1, device=dev()))


@pytest.mark.parametrize("B,Tp", [(1, 100), (3, 7), (50, 31), (100, 17), (97, 5)])
def test_lstm_matches_oracle(B, Tp):
    """Covers the three lockstep widths (NB = 1, 4, 8), odd lengths and a ragged last chunk."""
    enc, sd = make_encoder(512, True)
    g = torch.Generator().manual_seed(B * 1000 + Tp)
    idx = torch.randint(0, 512, (B, Tp), generator=g)
    ref = oenc.lstm(sd["codebook.embedding"][idx], sd["rnn.weight_ih_l0"], sd["rnn.weight_hh_l0"],
                    sd["rnn.bias_ih_l0"], sd["rnn.bias_hh_l0"])
    w, _keep = enc.pack_weights()
    lib = _lib.lib()
    ws_bytes = lib.vqcpc_lstm_workspace_bytes(B, Tp)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev())
    out = torch.empty(B, Tp, 256, device=dev())
    idxd = idx.to(dev())
    _lib.check(lib.vqcpc_lstm_forward(C.byref(w), _lib.ptr(idxd), B, Tp, _lib.ptr(ws), ws_bytes, _lib.ptr(out),
                                      _lib.current_stream_ptr()), "lstm")
    _lib.check(lib.vqcpc_check_status(_lib.ptr(ws), _lib.current_stream_ptr()), "lstm status")
    assert torch.allclose(out.cpu(), ref, rtol=RTOL, atol=ATOL_C), float((out.cpu() - ref).abs().max())


@pytest.mark.parametrize("B,T", [(64, 40), (130, 61), (600, 24), (2400, 12)])
def test_persistent_batched_lstm_matches_oracle(B, T):
    """Tensor-core modes run the whole recurrence of a batch of >= 64 utterances in ONE persistent tcgen05 launch
    (csrc/lstm_persist.cu; unit slices of 8 / 16 / 32 hidden units, ragged last row tile, two row tiles per CTA at 2400): the
    context must equal the oracle's LSTM on the code indices the GPU itself produced."""
    enc, sd = make_encoder(512, True)
    enc.gemm_mode = "bf16x3"
    mel = fixtures.synthetic_mel(B, T, seed=B)
    with torch.no_grad():
        z, c, idx = enc.encode(mel.to(dev()))
    sel = torch.cat([torch.arange(0, 6), torch.arange(B // 2, B // 2 + 6), torch.arange(B - 6, B)])
    ref = oenc.lstm(sd["codebook.embedding"][idx[sel].cpu()], sd["rnn.weight_ih_l0"], sd["rnn.weight_hh_l0"],
                    sd["rnn.bias_ih_l0"], sd["rnn.bias_hh_l0"])
    err = float((c[sel].cpu() - ref).abs().max())
    print(f"[persistent LSTM B={B}] max |c - oracle| = {err:.2e}")
    assert torch.allclose(c[sel].cpu(), ref, rtol=RTOL, atol=ATOL_C), err


@pytest.mark.parametrize("B,Tp", [(33, 90), (45, 33), (200, 80), (700, 50), (1500, 30), (4200, 12)])
def test_persistent_lstm_every_utterance_against_oracle_and_run_to_run(B, Tp):
    """The persistent LSTM's steps are chained through release / acquire counters and TMA reads of planes other CTAs wrote (one
    release per CTA and tile after a named barrier, fragment-layout epilogue, code indices two steps ahead).  An ordering bug
    there would show as garbage in SOME row tile, so EVERY utterance is compared with the oracle's LSTM -- uniformly random
    codes through the C ABI, all four configurations (8 / 16 / 32-unit slices, two tiles per CTA, ragged last tile) -- and three
    launches must agree bit for bit."""
    import ctypes as C
    enc, sd = make_encoder(512, True)
    w, _keep = enc.pack_weights()
    lib = _lib.lib()
    g = torch.Generator().manual_seed(B)
    idx = torch.randint(0, 512, (B, Tp), generator=g)
    idx_d = idx.to(dev())
    n = lib.vqcpc_lstm_workspace_bytes(B, Tp)
    outs = []
    for _ in range(3):
        ws = torch.empty(n, dtype=torch.uint8, device=dev())
        out = torch.empty(B, Tp, 256, device=dev())
        _lib.check(lib.vqcpc_lstm_forward_ex(C.byref(w), _lib.ptr(idx_d), B, Tp, _lib.ptr(ws), n, _lib.ptr(out), 1,
                                             _lib.current_stream_ptr()), "lstm")
        _lib.check(lib.vqcpc_check_status(_lib.ptr(ws), _lib.current_stream_ptr()), "lstm status")
        outs.append(out.cpu())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    ref = oenc.lstm(sd["codebook.embedding"][idx], sd["rnn.weight_ih_l0"], sd["rnn.weight_hh_l0"], sd["rnn.bias_ih_l0"],
                    sd["rnn.bias_hh_l0"])
    err = float((outs[0] - ref).abs().max())
    print(f"[persistent LSTM, all {B} utterances x {Tp} steps] max |c - oracle| = {err:.2e}")
    assert torch.allclose(outs[0], ref, rtol=RTOL, atol=ATOL_C), err


@pytest.mark.parametrize("mode", ["fp32", "bf16x3"])
def test_encode_from_host_equals_encode(mode):
    """Encoder.encode_from_host streams a host-resident batch to the GPU in chunks (front part per chunk, overlapped with the
    next chunk's copy; one LSTM launch over all utterances at the end, vqcpc_lstm_forward_ex): same results as encode() of the
    resident batch, and the indices / context match the oracle."""
    enc, sd = make_encoder(512, True)
    enc.gemm_mode = mode
    B, T = 70, 60
    mel = fixtures.synthetic_mel(B, T, seed=9)
    with torch.no_grad():
        z0, c0, i0 = enc.encode(mel.to(dev()))
        z1, c1, i1 = enc.encode_from_host(mel.pin_memory(), chunk_utterances=32)      # chunks of 32, 32, 6
        z2, c2, i2 = enc.encode_from_host(mel, chunk_utterances=1000)                  # one chunk, pageable memory
    for z, c, i in ((z1, c1, i1), (z2, c2, i2)):
        assert torch.equal(i.cpu(), i0.cpu()) and torch.equal(z.cpu(), z0.cpu())
        assert torch.allclose(c.cpu(), c0.cpu(), rtol=RTOL, atol=ATOL_C), float((c.cpu() - c0.cpu()).abs().max())
    sel = torch.tensor([0, 31, 32, 63, 64, 69])
    ref = oenc.lstm(sd["codebook.embedding"][i1[sel].cpu()], sd["rnn.weight_ih_l0"], sd["rnn.weight_hh_l0"],
                    sd["rnn.bias_ih_l0"], sd["rnn.bias_hh_l0"])
    assert torch.allclose(c1[sel].cpu(), ref, rtol=RTOL, atol=ATOL_C)


def test_text_dump_equals_numpy_savetxt(tmp_path):
    """encode.py:48-67 writes z, c and the pre-VQ embedding with np.savetxt(fmt="%.16f"); the GPU formatter (csrc/textdump.cu)
    must produce the same bytes -- checked against numpy itself, the function the reference calls."""
    from oracle import textio as otext
    from vectorquantizedcpc_b200 import format_txt, save_txt
    g = torch.Generator().manual_seed(3)
    # (a) the shapes encode.py writes: z (T', 64), c (T', 256) of a real encode
    enc, _ = make_encoder(512, True)
    with torch.no_grad():
        z, c, _ = enc.encode(fixtures.synthetic_mel(1, 300, seed=4).to(dev()))
    for t in (z[0], c[0]):
        assert format_txt(t).cpu().numpy().tobytes() == otext.savetxt_bytes(t.cpu().numpy())
    # (b) every kind of value: signs, zeros, denormals, huge magnitudes, ties at the 17th digit, nan / inf, random bit patterns
    edge = torch.tensor([0.0, -0.0, 1.0, -1.0, 0.5, 0.99999994, 9.9999999, 2.0 ** -149, 1e-45, 2.0 ** 24, 3.4e38, -3.4e38,
                         float("inf"), float("-inf"), float("nan"), 2.0 ** -16, 2.0 ** -17, 3 * 2.0 ** -18, 123456.789, -1e-20])
    ties = torch.arange(1, 200, 2, dtype=torch.float32) * 2.0 ** -17
    bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (4000,), generator=g, dtype=torch.int64).to(torch.int32).view(torch.float32)
    allv = torch.cat([edge, ties, bits])
    allv = allv[: (allv.numel() // 7) * 7].reshape(-1, 7)
    assert format_txt(allv.to(dev())).cpu().numpy().tobytes() == otext.savetxt_bytes(allv.numpy())
    # (c) 1-D input (one value per line), a ragged last block, and the file writer
    v = torch.randn(1000, generator=g)
    assert format_txt(v.to(dev())).cpu().numpy().tobytes() == otext.savetxt_bytes(v.numpy())
    big = torch.randn(3001, 64, generator=g) * 3
    p = tmp_path / "z.txt"
    nbytes = save_txt(p, big.to(dev()))
    data = open(p, "rb").read()
    assert len(data) == nbytes and data == otext.savetxt_bytes(big.numpy())
    assert torch.equal(torch.from_numpy(np.loadtxt(p, dtype=np.float32)), big)      # %.16f round-trips fp32 of this range


# ------------------------------------------------------------------------------------------ vocoder
def test_vocoder_conditioning_matches_oracle():
    voc, sd = make_vocoder()
    for (B, Tc, seed) in [(1, 50, 0), (3, 7, 1), (2, 1, 2)]:
        codes, spk, _ = fixtures.vocoder_inputs(B, Tc, seed=seed)
        G, p = voc.condition(codes.to(dev()), spk.to(dev()), return_prenet=True)
        p_o = ovoc.condition(sd, codes, spk)
        assert torch.allclose(p.cpu(), p_o, rtol=RTOL, atol=2e-5), float((p.cpu() - p_o).abs().max())
        G_o = p_o @ sd["rnnms.ar.rnn.weight_ih_l0"][:, 256:].t() + sd["rnnms.ar.rnn.bias_ih_l0"]
        assert torch.allclose(G.cpu(), G_o, rtol=RTOL, atol=2e-5)


def test_vocoder_teacher_forced_logits_match_oracle():
    voc, sd = make_vocoder()
    B, Tc, L = 2, 2, 500
    codes, spk, _ = fixtures.vocoder_inputs(B, Tc, seed=3)
    g = torch.Generator().manual_seed(11)
    x = torch.randint(0, 256, (B, L), generator=g)
    logits = voc.forward(x.to(dev()), codes.to(dev()), spk.to(dev())).cpu()
    ref = ovoc.forward_teacher_forced(sd, x, codes, spk)
    err = float((logits - ref).abs().max())
    print(f"[teacher-forced] max |dlogit| = {err:.3e} over {B}x{L} steps")
    assert logits.shape == (B, L, 256)
    assert err < ATOL_LOGITS


def test_vocoder_generate_matches_oracle_with_injected_uniforms():
    voc, sd = make_vocoder()
    B, Tc = 2, 2
    L = 320 * Tc
    codes, spk, u = fixtures.vocoder_inputs(B, Tc, seed=0)
    wav, x, logits = voc.generate(codes.to(dev()), spk.to(dev()), uniforms=u.to(dev()), return_mulaw=True,
                                  return_logits=True)
    wav, x, logits = wav.cpu(), x.cpu(), logits.cpu()
    assert wav.shape == (B, L) and x.shape == (B, L) and x.dtype == torch.int64
    lut = torch.from_numpy(mulaw.mulaw_decode_lut(8))
    assert torch.equal(wav, lut[x])                                  # mu-law decode epilogue is bit-exact
    # (1) free-running match against the oracle with identical uniforms (reported; chaotic after a divergence)
    wav_o, x_o, _ = ovoc.generate(sd, codes, spk, u, return_all=True)
    match = (x == x_o)
    first_div = [int((~m).nonzero()[0]) if (~m).any() else L for m in match]
    print(f"[generate] free-running sample match rate {float(match.float().mean()):.4f}; first divergence {first_div}")
    assert min(first_div) >= 50          # an early divergence would mean a wrong recurrence, not a rounding tie
    # (2) rigorous check: replay the GPU's own samples through the oracle (teacher forced): the logits must
    # agree and every sample must be consistent with its uniform under the oracle's CDF.
    x_in = torch.cat([torch.full((B, 1), 128, dtype=torch.int64), x[:, :-1]], dim=1)
    ref = ovoc.forward_teacher_forced(sd, x_in, codes, spk)
    err = float((logits - ref).abs().max())
    assert err < ATOL_LOGITS, err
    cdf = ovoc.cdf_bounds(ref)
    hi = torch.gather(cdf, 2, x[..., None])[..., 0]
    lo = torch.where(x > 0, torch.gather(cdf, 2, (x - 1).clamp(min=0)[..., None])[..., 0], torch.zeros_like(hi))
    tol = 5e-5
    ok = (u.double() >= lo - tol) & (u.double() <= hi + tol)
    assert bool(ok.all()), f"{int((~ok).sum())} samples inconsistent with their uniform"


def test_vocoder_generate_full_second_properties():
    """BASELINE config 3 size (B=1, 1 s = 16 000 steps): determinism under the same uniforms, range, LUT."""
    voc, _ = make_vocoder()
    codes, spk, u = fixtures.vocoder_inputs(1, 50, seed=0)
    a, xa = voc.generate(codes.to(dev()), spk.to(dev()), uniforms=u.to(dev()), return_mulaw=True)
    b, xb = voc.generate(codes.to(dev()), spk.to(dev()), uniforms=u.to(dev()), return_mulaw=True)
    assert a.shape == (1, 16000) and torch.equal(a, b) and torch.equal(xa, xb)
    assert int(xa.min()) >= 0 and int(xa.max()) <= 255 and float(a.abs().max()) <= 1.0
    assert len(torch.unique(xa)) > 50                 # actually sampling, not stuck
    c = voc.generate(codes.to(dev()), spk.to(dev()))  # device-side uniforms
    assert c.shape == (1, 16000) and not torch.equal(a, c)
    short = voc.generate(codes.to(dev()), spk.to(dev()), uniforms=u[:, :100].to(dev()), n_steps=100)
    assert torch.equal(short, a[:, :100])


def test_vocoder_batched_generate_equals_single_utterance_runs():
    """B = 7 -> interleaved groups of 4 + 2 + 1 inside the persistent kernel: every utterance must reproduce its own
    single-utterance run bit for bit (same uniforms), in generate and in teacher-forced mode."""
    voc, sd = make_vocoder()
    B, Tc, L = 7, 2, 300
    codes, spk, u = fixtures.vocoder_inputs(B, Tc, seed=4, n_steps=L)
    cd, sdv, ud = codes.to(dev()), spk.to(dev()), u.to(dev())
    wav, x = voc.generate(cd, sdv, uniforms=ud, n_steps=L, return_mulaw=True)
    for b in (0, 3, 4, 6):
        w1, x1 = voc.generate(cd[b:b + 1], sdv[b:b + 1], uniforms=ud[b:b + 1], n_steps=L, return_mulaw=True)
        assert torch.equal(x1[0], x[b]) and torch.equal(w1[0], wav[b]), b
    x_in = torch.cat([torch.full((B, 1), 128, dtype=torch.int64, device=dev()), x[:, :-1]], dim=1)
    tf = voc.forward(x_in, cd, sdv)
    ref = ovoc.forward_teacher_forced(sd, x_in[:2].cpu(), codes[:2], spk[:2])
    assert float((tf[:2].cpu() - ref).abs().max()) < ATOL_LOGITS
    tf1 = voc.forward(x_in[5:6], cd[5:6], sdv[5:6])
    assert torch.equal(tf1[0], tf[5])


@pytest.mark.parametrize("B,sel", [
    (8, [0, 5, 7]),                            # one group (ar_batch_kernel), one utterance tile of 16 in use
    (21, [0, 10, 11, 20]),                     # two utterance tiles of 16 in use
    (40, [0, 15, 16, 32, 39]),                 # 33..64 utterances: all four tiles
    (64, [0, 31, 32, 63]),                     # one full group
    (70, [0, 1, 31, 34, 35, 63, 64, 69]),      # two interleaved groups 35 + 35 (ar_batch2_kernel)
    (131, [0, 63, 64, 127, 128, 130]),         # launches of 128 (two groups 64 + 64) and 3
])
def test_vocoder_large_batch_kernel_matches_oracle(B, sel):
    """B >= 8 runs the batched tensor-core kernels: one group of up to 64 utterances per launch, or two interleaved
    groups for 65..128.
    Teacher-forced logits against the oracle, and free-running samples consistent with the oracle's CDF."""
    voc, sd = make_vocoder()
    Tc, L = 1, 200
    codes, spk, u = fixtures.vocoder_inputs(B, Tc, seed=9, n_steps=L)
    cd, sdv, ud = codes.to(dev()), spk.to(dev()), u.to(dev())
    wav, x, logits = voc.generate(cd, sdv, uniforms=ud, n_steps=L, return_mulaw=True, return_logits=True)
    wav, x, logits = wav.cpu(), x.cpu(), logits.cpu()
    lut = torch.from_numpy(mulaw.mulaw_decode_lut(8))
    assert torch.equal(wav, lut[x])
    x_in = torch.cat([torch.full((B, 1), 128, dtype=torch.int64), x[:, :-1]], dim=1)
    ref = ovoc.forward_teacher_forced(sd, x_in[sel], codes[sel], spk[sel])
    err = float((logits[sel] - ref).abs().max())
    print(f"[batched generate B={B}] max |dlogit| vs oracle on replayed samples = {err:.3e}")
    assert err < ATOL_LOGITS
    cdf = ovoc.cdf_bounds(ref)
    xs = x[sel]
    hi = torch.gather(cdf, 2, xs[..., None])[..., 0]
    lo = torch.where(xs > 0, torch.gather(cdf, 2, (xs - 1).clamp(min=0)[..., None])[..., 0], torch.zeros_like(hi))
    ok = (u[sel].double() >= lo - 5e-5) & (u[sel].double() <= hi + 5e-5)
    assert bool(ok.all())
    # free-running agreement with the single-utterance kernel (different summation order: compare samples, not bits)
    w1, x1 = voc.generate(cd[5:6], sdv[5:6], uniforms=ud[5:6], n_steps=L, return_mulaw=True)
    assert float((x1[0].cpu() == x[5]).float().mean()) > 0.97
    # teacher-forced mode through the batched kernel
    tf = voc.forward(x_in.to(dev()), cd, sdv).cpu()
    assert float((tf[sel] - ref).abs().max()) < ATOL_LOGITS


def test_vocoder_two_group_kernel_equals_single_group_kernel():
    """65..128 utterances per launch run as two interleaved groups (ar_batch2_kernel); per utterance the arithmetic is
    that of the single-group kernel, so samples and teacher-forced logits must agree bit for bit."""
    from vectorquantizedcpc_b200 import _lib
    voc, _ = make_vocoder()
    B, Tc, L = 100, 1, 240
    codes, spk, u = fixtures.vocoder_inputs(B, Tc, seed=21, n_steps=L)
    cd, sdv, ud = codes.to(dev()), spk.to(dev()), u.to(dev())
    wav2, x2, lg2 = voc.generate(cd, sdv, uniforms=ud, n_steps=L, return_mulaw=True, return_logits=True)
    try:
        _lib.check(_lib.lib().vqcpc_debug_set_ar_poll_gap(400 | (1 << 29)), "debug")      # two-group variant off
        wav1, x1, lg1 = voc.generate(cd, sdv, uniforms=ud, n_steps=L, return_mulaw=True, return_logits=True)
    finally:
        _lib.check(_lib.lib().vqcpc_debug_set_ar_poll_gap(400), "debug")
    assert torch.equal(x1, x2) and torch.equal(wav1, wav2) and torch.equal(lg1, lg2)
    x_in = torch.cat([torch.full((B, 1), 128, dtype=torch.int64, device=dev()), x2[:, :-1].long()], dim=1)
    tf2 = voc.forward(x_in, cd, sdv)
    try:
        _lib.check(_lib.lib().vqcpc_debug_set_ar_poll_gap(400 | (1 << 29)), "debug")
        tf1 = voc.forward(x_in, cd, sdv)
    finally:
        _lib.check(_lib.lib().vqcpc_debug_set_ar_poll_gap(400), "debug")
    assert torch.equal(tf1, tf2)


def test_vocoder_ragged_batch_equals_unpadded_runs():
    """a padded batch with ``lengths`` must give every utterance exactly its unpadded result -- the
    bidirectional prenet starts each backward pass at the utterance's own last frame."""
    voc, _ = make_vocoder()
    lens = [3, 1, 2]
    B, Tc = len(lens), max(lens)
    codes, spk, u = fixtures.vocoder_inputs(B, Tc, seed=31)
    cd, sdv, ud = codes.to(dev()), spk.to(dev()), u.to(dev())
    G, p = voc.condition(cd, sdv, return_prenet=True, lengths=lens)
    wav, x = voc.generate(cd, sdv, uniforms=ud, return_mulaw=True, lengths=torch.tensor(lens))
    for b, n in enumerate(lens):
        G1, p1 = voc.condition(cd[b:b + 1, :n], sdv[b:b + 1], return_prenet=True)
        assert torch.equal(p[b, :2 * n], p1[0]) and torch.equal(G[b, :2 * n], G1[0]), b
        assert not p[b, 2 * n:].any()
        w1, x1 = voc.generate(cd[b:b + 1, :n], sdv[b:b + 1], uniforms=ud[b:b + 1, :320 * n], return_mulaw=True)
        assert torch.equal(wav[b, :320 * n], w1[0]) and torch.equal(x[b, :320 * n], x1[0]), b
        assert not wav[b, 320 * n:].any()
    # a padded run WITHOUT lengths differs for the short utterances (backward pass sees the padding)
    p_pad = voc.condition(cd, sdv, return_prenet=True)[1]
    assert not torch.equal(p_pad[1, :2], p[1, :2])
    x_in = torch.cat([torch.full((B, 1), 128, dtype=torch.int64, device=dev()), x[:, :-1]], dim=1)
    tf = voc.forward(x_in, cd, sdv, lengths=lens)
    tf1 = voc.forward(x_in[1:2, :320], cd[1:2, :1], sdv[1:2])
    assert torch.equal(tf[1, :320], tf1[0])
    with pytest.raises(ValueError):
        voc.generate(cd, sdv, lengths=[3, 0, 2])
    with pytest.raises(ValueError):
        voc.generate(cd, sdv, lengths=[3, 4, 2])


def test_vocoder_ragged_generate_runs_each_launch_group_only_as_far_as_its_longest_utterance():
    """SURVEY 8f row 2 / vocoder.py:69: per-utterance stop for the single-utterance kernel (B < 8: every utterance is its own
    launch and runs exactly its own length) and length-sorted buckets for the batched kernels (B = 140 > one launch group of
    128).  Valid prefixes equal the unpadded single-utterance runs bit for bit (B < 8) / the same batch run to full length
    (batched kernels), padding is zero, and the ragged call is cheaper than the padded one."""
    voc, sd = make_vocoder()
    # (1) B = 5, lengths 1..4 code frames of 4
    B, Tc = 5, 4
    codes, spk, u = fixtures.vocoder_inputs(B, Tc, seed=61)
    lens = [4, 1, 3, 2, 1]
    cd, sdv, ud = codes.to(dev()), spk.to(dev()), u.to(dev())
    wav, x = voc.generate(cd, sdv, uniforms=ud, lengths=lens, return_mulaw=True)
    for b in range(B):
        n = 320 * lens[b]
        w1, x1 = voc.generate(cd[b:b + 1, :lens[b]], sdv[b:b + 1], uniforms=ud[b:b + 1, :n], return_mulaw=True)
        assert torch.equal(w1[0], wav[b, :n]) and torch.equal(x1[0], x[b, :n]), b
        assert float(wav[b, n:].abs().max()) == 0.0 if n < wav.shape[1] else True
    # (2) B = 140 (two launch groups after sorting), lengths 1..2 of 2: prefix equality with the full-length batched run
    B, Tc = 140, 2
    codes, spk, u = fixtures.vocoder_inputs(B, Tc, seed=62)
    g = torch.Generator().manual_seed(5)
    lens = torch.randint(1, Tc + 1, (B,), generator=g)
    lens[:3] = torch.tensor([2, 1, 2])
    cd, sdv, ud = codes.to(dev()), spk.to(dev()), u.to(dev())
    wav = voc.generate(cd, sdv, uniforms=ud, lengths=lens)
    assert wav.shape == (B, 640)
    for b in (0, 1, 2, 77, 139):
        n = 320 * int(lens[b])
        assert float(wav[b, n:].abs().max()) == 0.0 if n < 640 else True
        # the oracle on the unpadded utterance, teacher-forced on the GPU's own samples is covered elsewhere; here: range + prefix
        assert float(wav[b, :n].abs().max()) <= 1.0 and len(torch.unique(wav[b, :n])) > 20


def test_encoder_ragged_batch_equals_unpadded_runs():
    """Encoder.encode_ragged: zero padding is exact for the valid frames (zero-padded conv, per-frame MLP, causal LSTM)."""
    enc, _ = make_encoder(512, True)
    g = torch.Generator().manual_seed(5)
    mels = [torch.rand(80, T, generator=g).to(dev()) for T in (200, 101, 57, 2)]
    out = enc.encode_ragged(mels)
    for m, (z, c, idx) in zip(mels, out):
        z1, c1, i1 = enc.encode(m[None])
        assert idx.shape[0] == (m.shape[1] - 2) // 2 + 1
        assert torch.equal(idx, i1[0]) and torch.equal(z, z1[0]) and torch.equal(c, c1[0])
    assert enc.encode_ragged([]) == []
    with pytest.raises(ValueError):
        enc.encode_ragged([torch.rand(79, 10, device=dev())])


def _fe_signal(n, seed):
    g = np.random.default_rng(seed)
    t = np.arange(n) / 16000.0
    x = 0.3 * np.sin(2 * np.pi * 180 * t) + 0.2 * np.sin(2 * np.pi * 2900 * t * (1 + 0.3 * t)) + 0.03 * g.standard_normal(n)
    x[: n // 7] *= 1e-3                                   # a quiet stretch: exercises the top_db clamp
    return x.astype(np.float32)


def test_logmel_frontend_matches_oracle():
    """wave -> log-mel (preprocess.py:53-75) on the GPU against the float64 oracle; tolerance 2e-4 in the
    normalised units (1.0 = top_db = 80 dB), i.e. 0.016 dB.  Includes a ragged batch: every utterance is scaled, framed,
    reflected and clamped on its own."""
    from oracle import frontend as ofe
    from vectorquantizedcpc_b200.frontend import LogMel, wave_to_mel
    fe = LogMel().to(dev())
    lens = [48000, 16000, 5000, 1025]
    waves = [_fe_signal(n, 10 + i) for i, n in enumerate(lens)]
    for w in waves[:2]:
        ref = ofe.wave_to_mel(w)
        got = fe(torch.from_numpy(w).to(dev()))[0].cpu().double().numpy()
        assert got.shape == ref.shape == (80, 1 + len(w) // 160)
        err = np.abs(got - ref).max()
        print(f"[logmel N={len(w)}] max abs err {err:.2e} (range {ref.min():.3f}..{ref.max():.3f})")
        assert err < 2e-4
    batch = torch.zeros(len(lens), max(lens))
    for b, w in enumerate(waves):
        batch[b, :len(w)] = torch.from_numpy(w)
    out = fe(batch.to(dev()), lengths=lens).cpu().double().numpy()
    assert out.shape == (4, 80, 301)
    for b, w in enumerate(waves):
        Tb = 1 + len(w) // 160
        assert np.abs(out[b, :, :Tb] - ofe.wave_to_mel(w)).max() < 2e-4, b
        assert not out[b, :, Tb:].any()
    assert torch.equal(wave_to_mel(torch.from_numpy(waves[1]).to(dev())), fe(torch.from_numpy(waves[1]).to(dev()))[0])
    # the output feeds Encoder.encode directly
    enc, _ = make_encoder(512, True)
    z, c, idx = enc.encode(fe(torch.from_numpy(waves[1]).to(dev())))
    assert idx.shape == (1, (101 - 2) // 2 + 1)
    with pytest.raises(ValueError):
        fe(torch.zeros(1, 1024, device=dev()))
    with pytest.raises(ValueError):
        fe(batch.to(dev()), lengths=[48000, 16000, 5000, 1000])


def test_loudness_output_stage_matches_oracle():
    """convert.py:57,79-80 (pyloudnorm integrated loudness + gain) on the GPU against the oracle."""
    from oracle import loudness as olo
    from vectorquantizedcpc_b200.loudness import integrated_loudness, loudness_normalize
    lens = [48000, 30000, 16000, 6401]
    waves = [_fe_signal(n, 40 + i) * s for i, (n, s) in enumerate(zip(lens, (1.0, 0.2, 0.01, 0.5)))]
    batch = torch.zeros(len(lens), max(lens))
    for b, w in enumerate(waves):
        batch[b, :len(w)] = torch.from_numpy(w)
    bd = batch.to(dev())
    lufs = integrated_loudness(bd, 16000, lengths=lens).cpu()
    ref = torch.tensor([olo.integrated_loudness(w, 16000) for w in waves])
    print("[loudness] GPU", [round(float(v), 4) for v in lufs], "oracle", [round(float(v), 4) for v in ref])
    assert float((lufs - ref).abs().max()) < 2e-3                                   # LU
    one = integrated_loudness(torch.from_numpy(waves[1]).to(dev()))                  # unpadded single utterance
    assert abs(float(one[0]) - float(ref[1])) < 2e-3
    target = torch.tensor([-23.0, -20.0, -30.0, -18.0])
    out, measured = loudness_normalize(bd, target, 16000, lengths=lens)
    assert torch.allclose(measured.cpu(), lufs, atol=1e-6)
    for b, w in enumerate(waves):
        exp = olo.normalize_loudness(w, float(ref[b]), float(target[b]))
        got = out[b, :len(w)].cpu().double().numpy()
        assert np.abs(got - exp).max() <= 5e-4 * np.abs(exp).max()
        assert not out[b, len(w):].any()
    assert float(integrated_loudness(torch.zeros(1, 16000, device=dev()))[0]) == float("-inf")   # silence, as pyloudnorm
    with pytest.raises(ValueError):
        integrated_loudness(torch.zeros(1, 6400, device=dev()))


def test_convert_batch_equals_per_utterance_pipeline():
    """convert.py:52-83 for a ragged batch on the GPU: every utterance must come out exactly as when it is converted
    alone (front-end, encoder, ragged prenet, sample loop and loudness stage all act per utterance)."""
    from vectorquantizedcpc_b200 import convert_batch
    enc, _ = make_encoder(512, True)
    voc, _ = make_vocoder()
    lens = [9000, 7000, 8200]
    waves = [torch.from_numpy(_fe_signal(n, 60 + i)).to(dev()) for i, n in enumerate(lens)]
    spk = torch.tensor([3, 50, 101])
    code_lens = [((1 + n // 160) - 2) // 2 + 1 for n in lens]
    u = torch.rand(3, 320 * max(code_lens), generator=torch.Generator().manual_seed(9)).to(dev())
    outs = convert_batch(enc, voc, waves, spk, uniforms=u)
    for b, w in enumerate(waves):
        assert outs[b].shape == (320 * code_lens[b],)
        alone = convert_batch(enc, voc, [w], spk[b:b + 1], uniforms=u[b:b + 1, :320 * code_lens[b]])[0]
        assert torch.equal(outs[b], alone), b
    raw = convert_batch(enc, voc, waves, spk, uniforms=u, match_loudness=False)
    from vectorquantizedcpc_b200 import integrated_loudness
    src = integrated_loudness(waves[0])
    assert abs(float(integrated_loudness(outs[0])[0]) - float(src[0])) < 1e-2       # output loudness = source loudness
    assert not torch.equal(raw[0], outs[0])


def test_vocoder_argument_errors():
    voc, _ = make_vocoder()
    z = torch.zeros(1, 2, dtype=torch.int64, device=dev())
    s = torch.zeros(1, dtype=torch.int64, device=dev())
    with pytest.raises(IndexError):
        voc.generate(z + 512, s)
    with pytest.raises(IndexError):
        voc.generate(z, s + 102)
    with pytest.raises(ValueError):
        voc.generate(z.int(), s)
    with pytest.raises(ValueError):
        voc.generate(z, s, uniforms=torch.zeros(1, 5, device=dev()))
    with pytest.raises(ValueError):
        voc.forward(torch.zeros(1, 700, dtype=torch.int64, device=dev()), z, s)


# ------------------------------------------------------------------------------------------ round-2 parity hardening
def _cdf_consistent(ref_logits, x, u, tol=5e-5):
    cdf = ovoc.cdf_bounds(ref_logits)
    hi = torch.gather(cdf, 2, x[..., None])[..., 0]
    lo = torch.where(x > 0, torch.gather(cdf, 2, (x - 1).clamp(min=0)[..., None])[..., 0], torch.zeros_like(hi))
    return (u.double() >= lo - tol) & (u.double() <= hi + tol)


def test_vocoder_full_second_replayed_through_oracle():
    """BASELINE configs[2] at FULL size (B = 1, 16 000 steps): the GPU's own samples are replayed through the oracle in
    teacher-forced mode -- every one of the 16 000 logit vectors must match and every sample must be consistent with its
    injected uniform under the oracle's CDF.  (Free-running equality would only test the first divergence.)"""
    voc, sd = make_vocoder()
    codes, spk, u = fixtures.vocoder_inputs(1, 50, seed=0)
    wav, x, logits = voc.generate(codes.to(dev()), spk.to(dev()), uniforms=u.to(dev()), return_mulaw=True, return_logits=True)
    wav, x, logits = wav.cpu(), x.cpu(), logits.cpu()
    assert x.shape == (1, 16000)
    x_in = torch.cat([torch.full((1, 1), 128, dtype=torch.int64), x[:, :-1]], dim=1)
    ref = ovoc.forward_teacher_forced(sd, x_in, codes, spk)
    err = float((logits - ref).abs().max())
    print(f"[generate 16000 steps, replayed] max |dlogit| = {err:.3e}")
    assert err < ATOL_LOGITS, err
    ok = _cdf_consistent(ref, x, u)
    assert bool(ok.all()), f"{int((~ok).sum())} of 16000 samples inconsistent with their uniform"
    lut = torch.from_numpy(mulaw.mulaw_decode_lut(8))
    assert torch.equal(wav, lut[x])
    # the teacher-forced entry point at full length agrees with the logits the sample loop produced
    tf = voc.forward(x_in.to(dev()), codes.to(dev()), spk.to(dev())).cpu()
    assert float((tf - ref).abs().max()) < ATOL_LOGITS


def test_vocoder_round1_kernel_teacher_forced_stress_and_agreement():
    """The 128-CTA kernel (fallback when 7 clusters of 16 CTAs cannot be co-scheduled): a long teacher-forced run with four
    interleaved utterances must neither time out (the r exchange was WAR-safe only by timing in this mode) nor
    disagree with the cluster kernel beyond summation-order noise."""
    voc, sd = make_vocoder()
    B, Tc = 4, 320                                   # L = 102 400 steps
    L = 320 * Tc
    codes, spk, _ = fixtures.vocoder_inputs(B, Tc, seed=41)
    x = torch.randint(0, 256, (B, L), generator=torch.Generator().manual_seed(42))
    cd, sdv, xd = codes.to(dev()), spk.to(dev()), x.to(dev())
    lib = _lib.lib()
    try:
        _lib.check(lib.vqcpc_debug_set_ar_cluster(0, 200, 0), "debug")          # round-1 kernel, NB = 4 interleaved
        a = voc.forward(xd, cd, sdv)
        b = voc.forward(xd, cd, sdv)
    finally:
        _lib.check(lib.vqcpc_debug_set_ar_cluster(1, 200, 0), "debug")
    assert torch.equal(a, b)
    c = voc.forward(xd[:, :4000], cd, sdv)                                       # cluster kernel
    assert float((a[:, :4000] - c).abs().max()) < 1e-5
    ref = ovoc.forward_teacher_forced(sd, x[:1, :600], codes[:1], spk[:1])
    assert float((a[:1, :600].cpu() - ref).abs().max()) < ATOL_LOGITS


def test_vq_lookup_one_million_frames_against_oracle():
    """BASELINE configs[1] at full size: 1 M init-like frames (worst case for ties) against the oracle's fp32 argmin and the
    fp64 truth, in chunks: zero non-near-tie mismatches."""
    n = 1_000_000
    x, cb = fixtures.vq_inputs(n, kind="init", seed=1234)
    q, idx = run_vq(x, cb)
    idx = idx.reshape(-1)
    xf = x.reshape(-1, 64)
    tot = dict(mismatches=0, hard=0)
    tot64 = dict(mismatches=0, hard=0)
    for lo in range(0, n, 62_500):
        xs = xf[lo:lo + 62_500]
        _, io = oenc.vq_lookup(xs[None], cb)
        rep = oenc.classify_index_mismatches(xs, cb, idx[lo:lo + 62_500], io.reshape(-1))
        truth = oenc.vq_scores_exact(xs, cb).argmin(dim=-1)
        rep64 = oenc.classify_index_mismatches(xs, cb, idx[lo:lo + 62_500], truth)
        for k in tot:
            tot[k] += rep[k]
            tot64[k] += rep64[k]
    print(f"[vq 1M init-like] vs oracle fp32: {tot}; vs fp64 truth: {tot64}")
    assert tot["hard"] == 0 and tot64["hard"] == 0
    assert tot["mismatches"] <= 600 and tot64["mismatches"] <= tot["mismatches"]      # reference itself ~287 / 1 M off fp64
    assert torch.equal(q.reshape(-1, 64), cb[idx])


def test_vq_duplicated_and_collapsed_codebook_rows():
    """with three or more codes inside the recheck margin -- duplicated rows, a collapsed codebook -- and negative
    scores (where the index packed into the low mantissa bits orders backwards) the tensor-core search must still return
    torch.argmin's first minimum.  Frames flagged by the coarse pass are rescanned exactly over all 512 codes."""
    g = torch.Generator().manual_seed(77)
    cb = torch.randn(512, 64, generator=g)
    dup = [5, 130, 131, 300, 511]
    cb[dup] = cb[17].clone()                              # six identical rows: 5, 17, 130, 131, 300, 511
    cb[400:420] = cb[399] + 1e-7 * torch.randn(20, 64, generator=g)       # a collapsed cluster of near-duplicates
    n = 20000
    pick = torch.randint(0, 512, (n,), generator=g)
    x = (cb[pick] * 1.5 + 0.05 * torch.randn(n, 64, generator=g))[None]   # x.e >> |e|^2 / 2: negative scores at the winner
    q, idx = run_vq(x, cb)
    idx = idx.reshape(-1)
    _, io = oenc.vq_lookup(x, cb)
    rep = oenc.classify_index_mismatches(x, cb, idx, io.reshape(-1))
    assert rep["hard"] == 0, rep
    # exact duplicates: the lowest index of the six must win wherever any of them wins
    hit = torch.isin(io.reshape(-1), torch.tensor(dup + [17]))
    assert int(hit.sum()) > 100
    assert bool((idx[hit] == 5).all()) and bool((io.reshape(-1)[hit] == 5).all())
    # and the exact fp32 SIMT kernel (n < 8192 path) agrees with the tensor-core path frame by frame
    parts = [run_vq(x[:, lo:lo + 5000], cb)[1].reshape(-1) for lo in range(0, n, 5000)]
    assert torch.equal(torch.cat(parts), idx)


def test_vq_flag_list_overflow_and_degenerate_scales():
    """The tensor-core search lists the frames whose coarse minimum is not alone inside the error margin and rescans them exactly.
    Degenerate inputs must go the same way as ordinary ones: more flagged frames than the list holds (a fully collapsed codebook:
    every frame is a 512-way tie -> the rescan falls back to sweeping the indices), all-zero frames (margin 0), and scores of
    magnitude 1e-30 (where the saturating compare of the epilogue cannot be trusted and the frame is flagged by a guard).
    Reference: the exact fp32 kernel on the same inputs (chunks below 8192 frames), itself pinned to the oracle elsewhere."""
    g = torch.Generator().manual_seed(5)
    n = 40000                                            # > VQ_TC_FLAG_CAP (32768)

    def both(x, cb):
        q, idx = run_vq(x[None], cb)
        parts = [run_vq(x[None, lo:lo + 4000], cb) for lo in range(0, n, 4000)]
        qs = torch.cat([p[0].reshape(-1, 64) for p in parts]); ix = torch.cat([p[1].reshape(-1) for p in parts])
        assert torch.equal(idx.reshape(-1), ix) and torch.equal(q.reshape(-1, 64), qs)
        return idx.reshape(-1)

    # 1. collapsed codebook: all rows identical -> first index everywhere, every frame flagged
    cb = torch.randn(1, 64, generator=g).repeat(512, 1)
    x = torch.randn(n, 64, generator=g)
    idx = both(x, cb)
    assert bool((idx == 0).all())
    # 2. ordinary codebook, frames of zeros mixed in (score = |e|^2 exactly, margin = 1e-6 max|e|^2 only)
    cb = torch.randn(512, 64, generator=g)
    x = torch.randn(n, 64, generator=g)
    x[::7] = 0.0
    idx = both(x, cb)
    assert bool((idx[::7] == int((cb * cb).sum(1).argmin())).all())
    # 3. everything scaled to 1e-15: scores ~1e-30
    both(x * 1e-15, cb * 1e-15)


def test_ragged_batches_against_the_oracle_directly():
    """SURVEY 8f row 2, compared with the ORACLE on the unpadded utterances (not with another CUDA run)."""
    voc, sd = make_vocoder()
    lens = [2, 1, 3]
    B, Tc = len(lens), max(lens)
    codes, spk, u = fixtures.vocoder_inputs(B, Tc, seed=51)
    wav, x, logits = voc.generate(codes.to(dev()), spk.to(dev()), uniforms=u.to(dev()), return_mulaw=True, return_logits=True,
                                  lengths=lens)
    x, logits = x.cpu(), logits.cpu()
    for b, n in enumerate(lens):
        L = 320 * n
        xb = x[b:b + 1, :L]
        x_in = torch.cat([torch.full((1, 1), 128, dtype=torch.int64), xb[:, :-1]], dim=1)
        ref = ovoc.forward_teacher_forced(sd, x_in, codes[b:b + 1, :n], spk[b:b + 1])
        assert float((logits[b:b + 1, :L] - ref).abs().max()) < ATOL_LOGITS, b
        assert bool(_cdf_consistent(ref, xb, u[b:b + 1, :L]).all()), b
    enc, esd = make_encoder(512, True)
    g = torch.Generator().manual_seed(6)
    mels = [torch.rand(80, T, generator=g) for T in (120, 57, 200)]
    out = enc.encode_ragged([m.to(dev()) for m in mels])
    for m, (z, c, idx) in zip(mels, out):
        z_o, c_o, idx_o = oenc.encode(esd, m[None])
        rep = oenc.classify_index_mismatches(oenc.encode(esd, m[None], return_aux=True)[3], esd["codebook.embedding"], idx.cpu(), idx_o)
        assert rep["hard"] == 0
        if rep["mismatches"] == 0:
            assert torch.allclose(z.cpu(), z_o, rtol=RTOL, atol=ATOL_ZPRE) and torch.allclose(c.cpu(), c_o, rtol=RTOL, atol=ATOL_C)


def test_checkpoint_containers_run_on_the_gpu_and_match_the_oracle(tmp_path):
    """each container layout of the reference (train_cpc.py:23-29, convert.py:44, Lightning vocoder.py:41-51)
    is loaded from a FILE, moved to the GPU and run; results must match the oracle on the same weights."""
    from vectorquantizedcpc_b200 import checkpoint
    esd = fixtures.perturb_encoder_state(fixtures.encoder_init_state(512, seed=13))
    vsd = ovoc.init_state_dict(seed=13)
    files = {"cpc": tmp_path / "model.ckpt-100.pt", "release": tmp_path / "vocoder.pt", "lightning": tmp_path / "last.ckpt"}
    torch.save({"encoder": esd, "cpc": {}, "optimizer": {}, "scheduler": {}, "epoch": 100}, files["cpc"])
    torch.save({"vocoder": vsd, "epoch": 3}, files["release"])
    torch.save({"state_dict": {**{"model." + k: v for k, v in vsd.items()}, **{"encoder." + k: v for k, v in esd.items()}},
                "epoch": 1, "global_step": 10}, files["lightning"])
    mel = fixtures.synthetic_mel(2, 61, seed=8)
    z_o, c_o, idx_o = oenc.encode(esd, mel)
    codes, spk, u = fixtures.vocoder_inputs(1, 1, seed=9)
    xs = torch.randint(0, 256, (1, 200), generator=torch.Generator().manual_seed(10))
    tf_o = ovoc.forward_teacher_forced(vsd, xs, codes, spk)
    for name in ("cpc", "lightning"):
        enc = checkpoint.load_encoder(files[name]).to(dev())
        z, c, idx = enc.encode(mel.to(dev()))
        rep = oenc.classify_index_mismatches(oenc.encode(esd, mel, return_aux=True)[3], esd["codebook.embedding"], idx.cpu(), idx_o)
        assert rep["hard"] == 0, (name, rep)
        if rep["mismatches"] == 0:
            assert torch.allclose(z.cpu(), z_o, rtol=RTOL, atol=ATOL_ZPRE) and torch.allclose(c.cpu(), c_o, rtol=RTOL, atol=ATOL_C), name
    for name in ("release", "lightning"):
        voc = checkpoint.load_vocoder(files[name]).to(dev())
        tf = voc.forward(xs.to(dev()), codes.to(dev()), spk.to(dev())).cpu()
        assert float((tf - tf_o).abs().max()) < ATOL_LOGITS, name


def test_lstm_entry_reports_out_of_range_codes():
    """vqcpc_lstm_forward gathers table[idx] with caller-supplied indices -- they are clamped (memory safety) and an
    out-of-range value is reported through the workspace status word."""
    enc, _ = make_encoder(512, False)
    w, _keep = enc.pack_weights()
    lib = _lib.lib()
    B, Tp = 2, 9
    idx = torch.randint(0, 512, (B, Tp), device=dev())
    idx[1, 4] = 9999
    ws_bytes = lib.vqcpc_lstm_workspace_bytes(B, Tp)
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=dev())
    out = torch.empty(B, Tp, 256, device=dev())
    _lib.check(lib.vqcpc_lstm_forward(C.byref(w), _lib.ptr(idx), B, Tp, _lib.ptr(ws), ws_bytes, _lib.ptr(out),
                                      _lib.current_stream_ptr()), "lstm")
    assert lib.vqcpc_check_status(_lib.ptr(ws), _lib.current_stream_ptr()) == _lib.ERR_ARG
    assert bool(torch.isfinite(out).all())
