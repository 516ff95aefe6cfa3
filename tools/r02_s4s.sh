#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "encode or encoder or smoke or bf16" > gpurun_out/pytest_s4s.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_s4s.log
python tools/encode_profile.py 4096 bf16x3 3 > gpurun_out/enc_s4s.log 2>&1
python tools/encode_profile.py 512 bf16x3 3 >> gpurun_out/enc_s4s.log 2>&1
python tools/encode_profile.py 4096 bf16 3 >> gpurun_out/enc_s4s.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:im2col --csv --log-file gpurun_out/im2col.csv python tools/encode_profile.py 4096 bf16x3 1 > /dev/null 2>&1
