#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "vocoder or convert" > gpurun_out/pytest_voc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_voc.log
python tools/cl_trace.py v6c 300 0 1 > gpurun_out/cl_trace_v6c.log 2>&1
for dm in "100 200" "150 150" "200 150" "50 125" "200 100" "250 0" "350 0"; do set -- $dm; python tools/cl_trace.py sw_$1_$2 $1 $2 1 2>&1 | head -1 >> gpurun_out/cl_trace_v6c.log; done
