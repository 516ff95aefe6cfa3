#!/bin/bash
mkdir -p gpurun_out
python tools/lstm_time.py 64 512 1024 2048 4096 > gpurun_out/lp9.log 2>&1
VQCPC_LP_EW=1 python tools/lstm_time.py 1024 2048 >> gpurun_out/lp9.log 2>&1
python -m pytest tests -m gpu -x -q >> gpurun_out/lp9.log 2>&1
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
