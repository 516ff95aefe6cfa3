#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/lpx.log
for d in 0 256; do VQCPC_LP_DEBUG=$d python tools/lstm_time.py 64 512 1024 2048 4096 >> gpurun_out/lpx.log 2>&1; done
VQCPC_LP_DEBUG=288 python tools/lstm_time.py 512 2>&1 | tail -3 >> gpurun_out/lpx.log
