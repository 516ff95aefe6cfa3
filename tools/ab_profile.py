"""batched sample loop under ncu: python tools/ab_profile.py [B] [code_frames]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures, vocoder as ovoc
from vectorquantizedcpc_b200 import Vocoder, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
Tc = int(sys.argv[2]) if len(sys.argv) > 2 else 10
flag = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda:0")
voc = Vocoder(); voc.load_state_dict(ovoc.init_state_dict(seed=13)); voc = voc.to(dev).eval()
codes, spk, u = fixtures.vocoder_inputs(B, Tc, seed=0)
cd, sd, ud = codes.to(dev), spk.to(dev), u.to(dev)
_lib.check(_lib.lib().vqcpc_debug_set_ar_poll_gap(400 | flag), 'dbg')
with torch.no_grad():
    for _ in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); wav = voc.generate(cd, sd, uniforms=ud); b.record(); torch.cuda.synchronize()
        L = wav.shape[1]
        print(f"B={B} L={L}: {a.elapsed_time(b):.2f} ms, {1e3 * a.elapsed_time(b) / L:.2f} us/step, {B * L / 16000 / (a.elapsed_time(b) * 1e-3):.0f}x real time aggregate")
