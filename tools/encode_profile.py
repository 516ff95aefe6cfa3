"""one batched encode (cfg4 shape or a share of it) for ncu launch lists: python tools/encode_profile.py B mode"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures
from vectorquantizedcpc_b200 import Encoder, ConfEncoder
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
dev = torch.device("cuda:0")
enc = Encoder(ConfEncoder(channels=768)); enc.load_state_dict(fixtures.encoder_init_state(768, 13)); enc = enc.to(dev).eval()
enc.gemm_mode = mode
mel = fixtures.synthetic_mel(B, 300, seed=0).to(dev)
with torch.no_grad():
    for _ in range(2):
        enc.encode(mel)
torch.cuda.synchronize()
print("done")
