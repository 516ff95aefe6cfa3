#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "encode or encoder or lstm or smoke or linear or layernorm or checkpoint or ragged or convert" > gpurun_out/pytest_s4w.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_s4w.log
bash tools/r02_s4v.sh
bash tools/r02_enc1.sh
