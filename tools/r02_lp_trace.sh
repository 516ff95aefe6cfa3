#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/lp10.log
for B in 512 1024 2048 4096; do VQCPC_LP_DEBUG=32 python tools/lstm_time.py $B 2>&1 | tail -3 >> gpurun_out/lp10.log; done
for d in 4 16 2 1 23; do VQCPC_LP_DEBUG=$d python tools/lstm_time.py 512 2048 4096 >> gpurun_out/lp10.log 2>&1; done
