#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "vq" > gpurun_out/pytest_vq.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_vq.log
python tools/vq_profile.py 1000000 init > gpurun_out/vq_time.log 2>&1
python tools/vq_profile.py 1000000 trained >> gpurun_out/vq_time.log 2>&1
python tools/vq_exactness.py 4000000 >> gpurun_out/vq_time.log 2>&1
