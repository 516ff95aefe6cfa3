#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/lp_ew.log
for ew in 1 2 4; do echo "== VQCPC_LP_EW=$ew" >> gpurun_out/lp_ew.log; VQCPC_LP_EW=$ew python tools/lstm_time.py 64 256 512 1024 2048 4096 >> gpurun_out/lp_ew.log 2>&1; done
echo "== all-lane fence (dbg 64)" >> gpurun_out/lp_ew.log
VQCPC_LP_DEBUG=64 python tools/lstm_time.py 512 2048 4096 >> gpurun_out/lp_ew.log 2>&1
echo "== trace" >> gpurun_out/lp_ew.log
for B in 512 2048 4096; do VQCPC_LP_DEBUG=32 python tools/lstm_time.py $B 2>&1 | tail -3 >> gpurun_out/lp_ew.log; done
python tools/lstm_check.py 64 100 512 1100 2304 4096 >> gpurun_out/lp_ew.log 2>&1
python -m pytest tests -m gpu -x -q -k "lstm or encoder" >> gpurun_out/lp_ew.log 2>&1
