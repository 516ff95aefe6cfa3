#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/lp_time.log
for d in ${LP_DBG:-0 1 2 16 19}; do VQCPC_LP_DEBUG=$d python tools/lstm_time.py 512 4096 >> gpurun_out/lp_time.log 2>&1; done
python -m pytest tests -m gpu -x -q -k "lstm or encoder" >> gpurun_out/lp_time.log 2>&1
python tools/lstm_check.py 512 4096 >> gpurun_out/lp_time.log 2>&1
