#!/bin/bash
mkdir -p gpurun_out
python tools/lstm_time.py 64 256 512 1024 2048 4096 > gpurun_out/lp9.log 2>&1
VQCPC_LP_EW=1 python tools/lstm_time.py 64 256 512 >> gpurun_out/lp9.log 2>&1
timeout 300 python tools/lstm_check.py 64 100 512 1100 4096 >> gpurun_out/lp9.log 2>&1
python -m pytest tests -m gpu -x -q -k "lstm or encoder or ragged" >> gpurun_out/lp9.log 2>&1
