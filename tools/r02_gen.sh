#!/bin/bash
# generate-path check: vocoder parity tests, trace, timing
mkdir -p gpurun_out
tag=${1:-v5}
python -m pytest tests -m gpu -x -q -k "vocoder or convert" > gpurun_out/pytest_voc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_voc.log
python tools/cl_trace.py $tag ${2:-200} 0 1 > gpurun_out/cl_trace_$tag.log 2>&1
python tools/gen_profile.py 1 50 3 >> gpurun_out/cl_trace_$tag.log 2>&1
