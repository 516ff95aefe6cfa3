#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "encoder or vq or lstm or linear or layernorm or smoke or ragged or checkpoint" > gpurun_out/pytest_enc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_enc.log
bash tools/r02_enc1.sh
