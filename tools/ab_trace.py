"""phase trace of the batched sample loop: python tools/ab_trace.py [B]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures, vocoder as ovoc
from vectorquantizedcpc_b200 import Vocoder, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
flag = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device("cuda:0")
voc = Vocoder(); voc.load_state_dict(ovoc.init_state_dict(seed=13)); voc = voc.to(dev).eval()
codes, spk, u = fixtures.vocoder_inputs(B, 10, seed=0)
lib = _lib.lib()
lib.vqcpc_debug_set_ar_poll_gap(400 | flag)
cd, sd, ud = codes.to(dev), spk.to(dev), u.to(dev)
n, t0 = 128, 1000
names = ["G(gates)", "barrier1", "P2a(fc1)", "signal2+slice1+wait2", "P3(fc2)", "logits hop (LL)", "P4(sample)", "codes hop (LL)+reduce"]
with torch.no_grad():
    voc.generate(cd, sd, uniforms=ud)
    for cta in (0, 77):
        buf = torch.zeros(n + 1, 8, dtype=torch.int64, device=dev)
        lib.vqcpc_debug_set_ar_trace(buf.data_ptr(), cta, t0, n + 1)
        voc.generate(cd, sd, uniforms=ud)
        lib.vqcpc_debug_set_ar_trace(None, 0, 0, 0)
        ts = buf.cpu().double()
        d = torch.empty(n, 8)
        for k in range(7): d[:, k] = ts[:n, k + 1] - ts[:n, k]
        d[:, 7] = ts[1:n + 1, 0] - ts[:n, 7]
        step = ts[1:n + 1, 0] - ts[:n, 0]
        print(f"cta {cta}: step median {float(step.median()):.0f} cycles; phases median:", {nm: round(float(d[:, i].median())) for i, nm in enumerate(names)})
