#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/lpx.log
python tools/lstm_time.py 64 512 1024 2048 4096 >> gpurun_out/lpx.log 2>&1
VQCPC_LSTM_CLUSTER=4 python tools/lstm_time.py 64 512 1024 2048 4096 >> gpurun_out/lpx.log 2>&1
VQCPC_LSTM_CLUSTER=4 VQCPC_LP_DEBUG=32 python tools/lstm_time.py 512 2>&1 | tail -3 >> gpurun_out/lpx.log
python bench.py --no-cpu > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
