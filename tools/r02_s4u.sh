#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do python tools/lstm1_time.py; done > gpurun_out/lstm1.log 2>&1
bash tools/r02_enc1.sh
