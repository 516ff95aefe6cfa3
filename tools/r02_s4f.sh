#!/bin/bash
mkdir -p gpurun_out
VQCPC_VQ_TRACE=48 python tools/vq_flags.py 1000000 trained 2>&1 | tail -9 > gpurun_out/vq_trace.log
