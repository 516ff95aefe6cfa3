#!/bin/bash
mkdir -p gpurun_out
python tools/lstm_time.py 10 16 32 48 63 64 > gpurun_out/lp11.log 2>&1
