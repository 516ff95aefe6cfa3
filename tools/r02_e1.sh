#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/e1.log
for m in fp32 bf16x3 bf16; do python tools/encode_profile.py 1 $m 50 >> gpurun_out/e1.log 2>&1; done
for m in fp32 bf16x3; do python tools/encode_profile.py 8 $m 50 >> gpurun_out/e1.log 2>&1; done
