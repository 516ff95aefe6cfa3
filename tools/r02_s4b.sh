#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "vq" > gpurun_out/pytest_s4b.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_s4b.log
python tools/vq_profile.py 1000000 init > gpurun_out/vq_s4b.log 2>&1
python tools/vq_profile.py 1000000 trained >> gpurun_out/vq_s4b.log 2>&1
VQCPC_VQ_TRACE=48 python tools/vq_profile.py 1000000 init 2>&1 | tail -7 >> gpurun_out/vq_s4b.log
python tools/vq_exactness.py 4000000 >> gpurun_out/vq_s4b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vq_launches_init.csv python tools/vq_profile.py 1000000 init > gpurun_out/vq_ncu_init.log 2>&1
