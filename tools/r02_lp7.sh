#!/bin/bash
mkdir -p gpurun_out
python tools/lstm_time.py 64 512 1024 2048 4096 > gpurun_out/lp7.log 2>&1
for B in 512 2048; do VQCPC_LP_DEBUG=32 python tools/lstm_time.py $B 2>&1 | tail -3 >> gpurun_out/lp7.log; done
python tools/lstm_check.py 64 512 1100 4096 >> gpurun_out/lp7.log 2>&1
