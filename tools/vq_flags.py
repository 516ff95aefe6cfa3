"""Flagged-frame count of the tensor-core VQ search and time of its three kernels (debug): python tools/vq_flags.py [n] [kind]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures
from vectorquantizedcpc_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else "init"
dev = torch.device("cuda:0")
x, cb = fixtures.vq_inputs(n, kind=kind, seed=1234)
x = x[0].to(dev).contiguous(); cb = cb.to(dev).contiguous()
lib = _lib.lib()
wsb = lib.vqcpc_vq_workspace_bytes()
ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
q = torch.empty_like(x); idx = torch.empty(n, dtype=torch.int64, device=dev)
def run():
    st = lib.vqcpc_vq_lookup(_lib.ptr(x), _lib.ptr(cb), n, 512, 64, _lib.ptr(q), _lib.ptr(idx), _lib.ptr(ws), wsb, _lib.current_stream_ptr())
    assert st == 0, st
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): run()
b.record(); torch.cuda.synchronize()
off = 1024 + 512 * 128 * 2 + 512 * 4 + 512 * 32
cnt = ws[off:off + 4].view(torch.int32).item()
print(f"{kind}: {n} frames, flagged {cnt} ({100.0 * cnt / n:.4f} %), {a.elapsed_time(b) / 20 * 1000:.1f} us per call (C ABI, back to back)")
