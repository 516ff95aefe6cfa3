#!/bin/bash
mkdir -p gpurun_out
python - > gpurun_out/enc1_plain.log 2>&1 <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
from oracle import fixtures
from vectorquantizedcpc_b200 import Encoder, ConfEncoder
dev = torch.device("cuda:0")
enc = Encoder(ConfEncoder(channels=768)); enc.load_state_dict(fixtures.encoder_init_state(768, 13)); enc = enc.to(dev).eval()
mel = fixtures.synthetic_mel(1, 200, seed=0).to(dev)
with torch.no_grad():
    for _ in range(5): enc.encode(mel)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): enc.encode(mel)
    b.record(); torch.cuda.synchronize()
print(f"encode 1 x 2 s fp32: {a.elapsed_time(b) / 50:.4f} ms")
PY
cat > /tmp/enc1.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
from oracle import fixtures
from vectorquantizedcpc_b200 import Encoder, ConfEncoder
dev = torch.device("cuda:0")
enc = Encoder(ConfEncoder(channels=768)); enc.load_state_dict(fixtures.encoder_init_state(768, 13)); enc = enc.to(dev).eval()
mel = fixtures.synthetic_mel(1, 200, seed=0).to(dev)
with torch.no_grad():
    for _ in range(3): enc.encode(mel)
torch.cuda.synchronize()
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/enc1_launches.csv python /tmp/enc1.py > gpurun_out/enc1_ncu.log 2>&1
