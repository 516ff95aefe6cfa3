#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/enc_launches_4096_bf16.csv python tools/encode_profile.py 4096 bf16 1 > gpurun_out/enc_ncu_4096_bf16.log 2>&1
