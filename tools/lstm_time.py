"""Time of the batched LSTM alone (vqcpc_lstm_forward_ex, tensor-core mode) on random code indices.
python tools/lstm_time.py [B ...]     VQCPC_LP_DEBUG=<bits> ablations of lstm_persist_kernel (timing only)."""
import os, sys, ctypes as C, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import fixtures
from vectorquantizedcpc_b200 import Encoder, ConfEncoder, _lib
Bs = [int(a) for a in sys.argv[1:]] or [512, 2048, 4096]
dev = torch.device("cuda:0")
sd = fixtures.perturb_encoder_state(fixtures.encoder_init_state(768, 13))
enc = Encoder(ConfEncoder(channels=768)); enc.load_state_dict(sd); enc = enc.to(dev).eval()
lib = _lib.lib()
Tp = 150
for B in Bs:
    w, keep = enc.pack_weights()
    idx = torch.randint(0, 512, (B, Tp), device=dev)
    c = torch.empty(B, Tp, 256, device=dev)
    n = lib.vqcpc_lstm_workspace_bytes(B, Tp)
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    def run():
        _lib.check(lib.vqcpc_lstm_forward_ex(C.byref(w), _lib.ptr(idx), B, Tp, _lib.ptr(ws), n, _lib.ptr(c), int(os.environ.get("LP_MODE", "1")), _lib.current_stream_ptr()), "lstm")
    run(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): run()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"B={B:5d} Tp={Tp}: LSTM {ms:.3f} ms = {ms * 1e3 / (Tp - 1):.2f} us/step  (VQCPC_LP_DEBUG={os.environ.get('VQCPC_LP_DEBUG', '0')})", flush=True)
