#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3 4 7; do echo "VQCPC_LC_DEBUG=$d"; VQCPC_LC_DEBUG=$d python tools/lstm1_time.py; done > gpurun_out/lstm1.log 2>&1
echo "L2 kernel"; VQCPC_LSTM_CLUSTER16=0 python tools/lstm1_time.py >> gpurun_out/lstm1.log 2>&1
