#!/bin/bash
mkdir -p gpurun_out
python tools/vq_flags.py 1000000 init > gpurun_out/vq_flags.log 2>&1
python tools/vq_flags.py 1000000 trained >> gpurun_out/vq_flags.log 2>&1
VQCPC_VQ_TRACE=48 python tools/vq_flags.py 1000000 trained 2>&1 | tail -9 >> gpurun_out/vq_flags.log
