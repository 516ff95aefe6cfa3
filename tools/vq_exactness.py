"""tensor-core VQ path vs the exact fp32 kernel at full size: python tools/vq_exactness.py [n_frames]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures
from vectorquantizedcpc_b200 import VQEmbeddingEMA
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
dev = torch.device("cuda:0")
for kind in ("init", "trained"):
    for seed in (1234, 99):
        x, cb = fixtures.vq_inputs(n, kind=kind, seed=seed)
        vq = VQEmbeddingEMA(512, 64); vq.embedding.copy_(cb); vq = vq.to(dev)
        xd = x.to(dev)
        q, idx = vq.encode(xd)                                   # tensor-core path (n >= 8192)
        flat = xd.reshape(-1, 64)
        ref = torch.cat([vq.encode(flat[i:i + 8000][None])[1].reshape(-1) for i in range(0, flat.shape[0], 8000)])   # fp32 kernel
        bad = (idx.reshape(-1) != ref).nonzero().flatten()
        print(f"{kind} seed {seed}: {flat.shape[0]} frames, mismatches tensor-core vs fp32 kernel: {bad.numel()}")
        if bad.numel():
            xb = flat[bad[:5]].double(); e = vq.embedding.double()
            d = (e * e).sum(1)[None] - 2 * xb @ e.T
            srt = d.sort(dim=1).values
            print("   fp64 gaps best->2nd, 2nd->3rd of the first mismatches:", (srt[:, 1] - srt[:, 0]).tolist(), (srt[:, 2] - srt[:, 1]).tolist())
