"""VQ lookup microbench (BASELINE configs[1]) for ncu / timing: python tools/vq_profile.py [n_frames] [kind]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures
from vectorquantizedcpc_b200 import VQEmbeddingEMA
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else "init"
dev = torch.device("cuda:0")
x, cb = fixtures.vq_inputs(n, kind=kind, seed=1234)
vq = VQEmbeddingEMA(512, 64); vq.embedding.copy_(cb); vq = vq.to(dev)
xd = x.to(dev)
for _ in range(3): vq.encode(xd)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): vq.encode(xd)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(f"vq_lookup {n} frames ({kind}): {ms:.4f} ms  {n / ms / 1e6:.2f} G frames/s  {520 * n / ms / 1e6:.0f} GB/s algorithmic")
