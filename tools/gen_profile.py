"""one Vocoder.generate call (B utterances x Tc code frames) with timing -- the command behind the ncu captures of the
sample-loop kernels: python tools/gen_profile.py [B] [code_frames] [reps]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures, vocoder as ovoc
from vectorquantizedcpc_b200 import Vocoder
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
Tc = int(sys.argv[2]) if len(sys.argv) > 2 else 50
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
voc = Vocoder(); voc.load_state_dict(ovoc.init_state_dict(seed=13)); voc = voc.to(dev).eval()
codes, spk, u = fixtures.vocoder_inputs(B, Tc, seed=0)
cd, sd, ud = codes.to(dev), spk.to(dev), u.to(dev)
with torch.no_grad():
    voc.generate(cd, sd, uniforms=ud)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        voc.generate(cd, sd, uniforms=ud)
    b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print(f"generate B={B} Tc={Tc}: {ms:.3f} ms per call, {ms * 1e3 / (320 * Tc):.3f} us/step, {B * 320 * Tc / 16000 / (ms * 1e-3):.1f}x RT aggregate")
