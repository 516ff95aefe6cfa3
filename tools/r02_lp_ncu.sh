#!/bin/bash
mkdir -p gpurun_out
bash tools/r02_enc_launches.sh
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_persist -s 1 -c 1 -f -o gpurun_out/r02_lstm_persist python tools/lstm_time.py 4096 > gpurun_out/ncu_lp.log 2>&1
