#!/bin/bash
mkdir -p gpurun_out
python tools/gen_profile.py 1 50 1 > gpurun_out/gen_plain.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ar_cluster -c 1 python tools/gen_profile.py 1 50 1 > gpurun_out/ncu_ar_a.log 2>&1; echo "a exit $?" >> gpurun_out/ncu_ar_a.log
timeout 400 ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section LaunchStats --section Occupancy --clock-control none -k regex:ar_cluster -c 1 -f -o gpurun_out/r02_ar_cluster_b python tools/gen_profile.py 1 50 1 > gpurun_out/ncu_ar_b.log 2>&1; echo "b exit $?" >> gpurun_out/ncu_ar_b.log
timeout 600 ncu --set full --import-source on --replay-mode application --clock-control none -k regex:ar_cluster -c 1 -f -o gpurun_out/r02_ar_cluster_c python tools/gen_profile.py 1 50 1 > gpurun_out/ncu_ar_c.log 2>&1; echo "c exit $?" >> gpurun_out/ncu_ar_c.log
