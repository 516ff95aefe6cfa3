"""Per-step time of the latency LSTM (one utterance) through the C ABI: python tools/lstm1_time.py   (VQCPC_LC_DEBUG ablations)"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures
from vectorquantizedcpc_b200 import Encoder, ConfEncoder, _lib
dev = torch.device("cuda:0")
enc = Encoder(ConfEncoder(channels=512)); enc.load_state_dict(fixtures.encoder_init_state(512, 13)); enc = enc.to(dev).eval()
w, _keep = enc.pack_weights()
lib = _lib.lib()
res = {}
for B in (1, 4):
    for Tp in (100, 1100):
        idx = torch.randint(0, 512, (B, Tp), device=dev)
        wsb = lib.vqcpc_lstm_workspace_bytes(B, Tp)
        ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
        out = torch.empty(B, Tp, 256, device=dev)
        def run():
            _lib.check(lib.vqcpc_lstm_forward(C.byref(w), _lib.ptr(idx), B, Tp, _lib.ptr(ws), wsb, _lib.ptr(out), _lib.current_stream_ptr()), "lstm")
        for _ in range(3): run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): run()
        b.record(); torch.cuda.synchronize()
        res[(B, Tp)] = a.elapsed_time(b) / 10 * 1000
    print(f"B={B}: {res[(B, 100)]:.1f} us at 100 steps, {res[(B, 1100)]:.1f} us at 1100 steps -> {(res[(B, 1100)] - res[(B, 100)]) / 1000:.3f} us/step, fixed {res[(B, 100)] - 100 * (res[(B, 1100)] - res[(B, 100)]) / 1000:.1f} us")
