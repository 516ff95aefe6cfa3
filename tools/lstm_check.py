"""Batched LSTM of Encoder.encode (tensor-core modes): context c against the oracle's LSTM run on the GPU's own code indices,
and the time of a whole encode.   python tools/lstm_check.py [B ...]      (VQCPC_LSTM_PERSIST=0: per-step-launch paths)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures, encoder as oenc
from vectorquantizedcpc_b200 import Encoder, ConfEncoder
Bs = [int(a) for a in sys.argv[1:]] or [64, 100, 512, 1100, 2304, 4096]
dev = torch.device("cuda:0")
sd = fixtures.perturb_encoder_state(fixtures.encoder_init_state(768, 13))
enc = Encoder(ConfEncoder(channels=768)); enc.load_state_dict(sd); enc = enc.to(dev).eval()
enc.gemm_mode = "bf16x3"
for B in Bs:
    T = 300 if B <= 4096 else 100
    mel = fixtures.synthetic_mel(B, T, seed=B).to(dev)
    with torch.no_grad():
        z, c, idx = enc.encode(mel)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): enc.encode(mel)
        b.record(); torch.cuda.synchronize()
    nchk = min(B, 24)
    sel = torch.cat([torch.arange(nchk // 2), torch.arange(B - nchk // 2, B)])
    ref = oenc.lstm(sd["codebook.embedding"][idx[sel].cpu()], sd["rnn.weight_ih_l0"], sd["rnn.weight_hh_l0"], sd["rnn.bias_ih_l0"], sd["rnn.bias_hh_l0"])
    err = float((c[sel].cpu() - ref).abs().max())
    print(f"B={B:5d} T={T}: encode {a.elapsed_time(b) / 3:.3f} ms   max |c - oracle LSTM(idx)| over {len(sel)} utterances = {err:.2e}", flush=True)
