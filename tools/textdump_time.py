"""Time of the GPU "%.16f" text dump against np.savetxt (encode.py:48-67): python tools/textdump_time.py"""
import io, os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vectorquantizedcpc_b200 import format_txt
dev = torch.device("cuda:0")
for rows, cols in ((100, 64), (150, 256), (15000, 64), (150000, 64)):
    x = torch.randn(rows, cols) * 2
    xd = x.to(dev)
    for _ in range(3): format_txt(xd)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): b = format_txt(xd).cpu()
    t_gpu = (time.perf_counter() - t0) / 10
    t0 = time.perf_counter(); buf = io.BytesIO(); np.savetxt(buf, x.numpy(), fmt="%.16f"); t_np = time.perf_counter() - t0
    assert b.numpy().tobytes() == buf.getvalue()
    print(f"{rows} x {cols}: GPU format + D2H {t_gpu * 1e3:.3f} ms ({b.numel() / t_gpu / 1e9:.2f} GB/s of text), np.savetxt {t_np * 1e3:.1f} ms ({t_np / t_gpu:.0f}x)")
