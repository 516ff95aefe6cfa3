#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/lp_time2.log
for d in 0 4 19 23; do VQCPC_LP_DEBUG=$d python tools/lstm_time.py 1024 2048 2304 4096 >> gpurun_out/lp_time2.log 2>&1; done
