#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/enc1.py <<'PY'
import sys, os, torch
sys.path.insert(0, os.getcwd())
from oracle import fixtures
from vectorquantizedcpc_b200 import Encoder, ConfEncoder
dev = torch.device("cuda:0")
enc = Encoder(ConfEncoder(channels=768)); enc.load_state_dict(fixtures.encoder_init_state(768, 13)); enc = enc.to(dev).eval()
mel = fixtures.synthetic_mel(1, 200, seed=0).to(dev)
with torch.no_grad():
    for _ in range(4): enc.encode(mel)
torch.cuda.synchronize()
PY
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/enc1_warm.csv python /tmp/enc1.py > gpurun_out/enc1_warm.log 2>&1
