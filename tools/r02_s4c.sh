#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:vq_tc_kernel -s 3 -c 1 -f -o gpurun_out/r02_vq_tc_v2 python tools/vq_profile.py 1000000 init > gpurun_out/ncu_vq.log 2>&1
