#!/bin/bash
mkdir -p gpurun_out
python tools/ab_trace.py 64 > gpurun_out/ab_trace_64.log 2>&1
python tools/ab_trace.py 16 >> gpurun_out/ab_trace_64.log 2>&1
