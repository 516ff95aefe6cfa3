"""one batched encode (cfg4 shape or a share of it), timed with CUDA events; also the command for ncu launch lists:
python tools/encode_profile.py B mode [reps]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures
from vectorquantizedcpc_b200 import Encoder, ConfEncoder
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda:0")
enc = Encoder(ConfEncoder(channels=768)); enc.load_state_dict(fixtures.encoder_init_state(768, 13)); enc = enc.to(dev).eval()
enc.gemm_mode = mode
mel = fixtures.synthetic_mel(B, 300, seed=0).to(dev)
with torch.no_grad():
    for _ in range(2):
        enc.encode(mel)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        enc.encode(mel)
    e1.record()
    torch.cuda.synchronize()
print(f"encode {B} x 3 s, mode {mode}: {e0.elapsed_time(e1) / reps:.3f} ms per pass")
