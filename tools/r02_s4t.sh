#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "encode or encoder or lstm or smoke or vq or linear or layernorm or checkpoint or ragged or convert" > gpurun_out/pytest_s4t.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_s4t.log
bash tools/r02_enc1.sh
python tools/lstm1_time.py > gpurun_out/lstm1.log 2>&1
