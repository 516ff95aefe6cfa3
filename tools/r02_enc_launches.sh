#!/bin/bash
mkdir -p gpurun_out
for B in 512 4096; do
python tools/encode_profile.py $B bf16x3 3 > gpurun_out/enc_plain_$B.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/enc_launches_$B.csv python tools/encode_profile.py $B bf16x3 1 > gpurun_out/enc_ncu_$B.log 2>&1
done
python tools/encode_profile.py 4096 bf16 3 >> gpurun_out/enc_plain_4096.log 2>&1
python tools/encode_profile.py 1 fp32 20 >> gpurun_out/enc_plain_4096.log 2>&1
