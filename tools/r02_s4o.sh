#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "lstm or encoder" > gpurun_out/pytest_s4o.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_s4o.log
python tools/lstm1_time.py > gpurun_out/lstm1.log 2>&1
for d in 1 2 3 4 7; do echo "VQCPC_LC_DEBUG=$d"; VQCPC_LC_DEBUG=$d python tools/lstm1_time.py | head -1; done >> gpurun_out/lstm1.log 2>&1
bash tools/r02_enc1.sh
