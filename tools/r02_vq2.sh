#!/bin/bash
mkdir -p gpurun_out
for kind in init trained; do
python tools/vq_profile.py 1000000 $kind > gpurun_out/vq_plain_$kind.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/vq_launches_$kind.csv python tools/vq_profile.py 1000000 $kind > gpurun_out/vq_ncu_$kind.log 2>&1
done
