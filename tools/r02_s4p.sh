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
    for _ in range(3): enc.encode(mel)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:sgemm_tn_kernel -s 8 -c 1 -f -o gpurun_out/r02_sgemm_small python /tmp/enc1.py > gpurun_out/ncu_sgemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_cluster -s 1 -c 1 -f -o gpurun_out/r02_lstm_cluster python /tmp/enc1.py > gpurun_out/ncu_lstmc.log 2>&1
