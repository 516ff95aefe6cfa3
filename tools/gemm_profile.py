"""a few big tcgen05 GEMMs (the encoder's C x C layer at cfg4 scale) for ncu: python tools/gemm_profile.py"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vectorquantizedcpc_b200 import _lib
dev = torch.device("cuda:0")
M, N, K = 614400, 768, 768
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) / K ** 0.5
out = torch.empty(M, N, device=dev)
ap = torch.empty(M, 2 * K, dtype=torch.bfloat16, device=dev); wp = torch.empty(N, 2 * K, dtype=torch.bfloat16, device=dev)
err = torch.zeros(1, dtype=torch.int32, device=dev)
lib = _lib.lib()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for i in range(4):
    if i == 1: ev[0].record()
    _lib.check(lib.vqcpc_linear_tc(A.data_ptr(), W.data_ptr(), None, out.data_ptr(), M, N, K, 3, ap.data_ptr(), wp.data_ptr(),
                                   err.data_ptr(), torch.cuda.current_stream().cuda_stream), "tc")
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 3
print(f"split + gemm_tc {M}x{N}x{K} (3 terms): {ms:.3f} ms per call, err flag {int(err)}")
