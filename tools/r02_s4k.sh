#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "vq" > gpurun_out/pytest_s4k.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_s4k.log
timeout 120 python tools/vq_flags.py 1000000 init > gpurun_out/vq_flags.log 2>&1
timeout 120 python tools/vq_flags.py 1000000 trained >> gpurun_out/vq_flags.log 2>&1
VQCPC_VQ_TRACE=48 timeout 120 python tools/vq_flags.py 1000000 trained 2>&1 | tail -9 >> gpurun_out/vq_flags.log
VQCPC_VQ_TRACE=48 timeout 120 python tools/vq_flags.py 1000000 init 2>&1 | tail -9 >> gpurun_out/vq_flags.log
timeout 200 python tools/vq_exactness.py 4000000 >> gpurun_out/vq_flags.log 2>&1
