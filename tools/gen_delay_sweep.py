"""us/step of Vocoder.generate B = 1 (production cluster kernel) against the first-probe delay of its grid hop.
python tools/gen_delay_sweep.py [delay ...]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures, vocoder as ovoc
from vectorquantizedcpc_b200 import Vocoder, _lib
delays = [int(a) for a in sys.argv[1:]] or [200, 225, 250, 275, 300, 325, 350, 400]
dev = torch.device("cuda:0")
voc = Vocoder(); voc.load_state_dict(ovoc.init_state_dict(seed=13)); voc = voc.to(dev).eval()
codes, spk, u = fixtures.vocoder_inputs(1, 50, seed=0)
cd, sdv, ud = codes.to(dev), spk.to(dev), u.to(dev)
lib = _lib.lib()
for rep in range(2):
    for d in delays:
        lib.vqcpc_debug_set_ar_cluster(1, d, 0)
        with torch.no_grad():
            voc.generate(cd, sdv, uniforms=ud); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(4): voc.generate(cd, sdv, uniforms=ud)
            b.record(); torch.cuda.synchronize()
        print(f"delay {d:4d}: {a.elapsed_time(b) / 4 / 16000 * 1e3:.4f} us/step", flush=True)
