#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/lp_trace.log
for d in 32 55; do for B in 512 2048 4096; do VQCPC_LP_DEBUG=$d python tools/lstm_time.py $B 2>&1 | tail -9 >> gpurun_out/lp_trace.log; done; done
VQCPC_LSTM_CLUSTER=4 python tools/lstm_time.py 512 2048 4096 >> gpurun_out/lp_trace.log 2>&1
