#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_ln_pair -s 7 -c 1 -f -o gpurun_out/r02_pair_bf16 python tools/encode_profile.py 4096 bf16 1 > gpurun_out/ncu_pair_bf16.log 2>&1
