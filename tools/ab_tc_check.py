"""tcgen05 batched sample loop (vocoder_batch_tc.cu) vs the two-group mma.sync kernel: python tools/ab_tc_check.py [B]"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures, vocoder as ovoc
from vectorquantizedcpc_b200 import Vocoder, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
voc = Vocoder(); voc.load_state_dict(ovoc.init_state_dict(seed=13)); voc = voc.to(dev).eval()
codes, spk, u = fixtures.vocoder_inputs(B, 2, seed=3)
cd, sd, ud = codes.to(dev), spk.to(dev), u.to(dev)
lib = _lib.lib()
def run(flag):
    _lib.check(lib.vqcpc_debug_set_ar_poll_gap(400 | flag), "dbg")
    try:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        voc.generate(cd, sd, uniforms=ud)
        a.record(); out = voc.generate(cd, sd, uniforms=ud, return_mulaw=True, return_logits=True); b.record(); torch.cuda.synchronize()
        return out, a.elapsed_time(b)
    finally:
        _lib.check(lib.vqcpc_debug_set_ar_poll_gap(400), "dbg")
(w0, x0, l0), t0 = run(0)
(w1, x1, l1), t1 = run(1 << 30)
L = w0.shape[1]
print(f"B={B} L={L}: two-group {t0:.2f} ms ({1e3*t0/L:.2f} us/step), tcgen05 {t1:.2f} ms ({1e3*t1/L:.2f} us/step, {B*L/16000/(t1*1e-3):.0f}x real time)")
same = (x0 == x1).float().mean().item()
first = (x0 != x1).float().argmax(dim=1)
print(f"sample agreement {same:.4f}; max |dlogit| at step 0: {(l0[:, 0] - l1[:, 0]).abs().max().item():.3e}; over agreeing prefix (first 50 steps): {(l0[:, :50] - l1[:, :50]).abs().max().item():.3e}")
x_in = torch.cat([torch.full((B, 1), 128, dtype=torch.int64, device=dev), x0[:, :-1]], dim=1)
tf0 = voc.forward(x_in, cd, sd)
_lib.check(lib.vqcpc_debug_set_ar_poll_gap(400 | (1 << 30)), "dbg")
tf1 = voc.forward(x_in, cd, sd)
_lib.check(lib.vqcpc_debug_set_ar_poll_gap(400), "dbg")
print(f"teacher-forced max |dlogit| tcgen05 vs two-group: {(tf0 - tf1).abs().max().item():.3e}")
ref = ovoc.forward_teacher_forced(ovoc.init_state_dict(seed=13), x_in[:2].cpu(), codes[:2], spk[:2])
print(f"teacher-forced max |dlogit| tcgen05 vs oracle (2 utterances): {(tf1[:2].cpu() - ref).abs().max().item():.3e}")
