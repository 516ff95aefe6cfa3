#!/bin/bash
# session 4, call a: new VQ epilogue (min + count) and the reworked cluster LSTM
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "vq or encoder or lstm" > gpurun_out/pytest_s4a.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_s4a.log
python tools/vq_profile.py 1000000 init > gpurun_out/vq_s4a.log 2>&1
python tools/vq_profile.py 1000000 trained >> gpurun_out/vq_s4a.log 2>&1
VQCPC_VQ_TRACE=48 python tools/vq_profile.py 1000000 init >> gpurun_out/vq_s4a.log 2>&1
python tools/vq_exactness.py 4000000 >> gpurun_out/vq_s4a.log 2>&1
bash tools/r02_enc1.sh
