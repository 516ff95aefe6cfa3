#!/bin/bash
mkdir -p gpurun_out
for g in 0 100 200 400 800 1500; do echo "VQCPC_AB_GAP=$g"; VQCPC_AB_GAP=$g python tools/ab_profile.py 64 10 | tail -1; VQCPC_AB_GAP=$g python tools/ab_profile.py 16 10 | tail -1; done > gpurun_out/ab_gap.log 2>&1
