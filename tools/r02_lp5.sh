#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/lp_ew2.log
for d in 32 33 34 48 36 51 55; do echo "== dbg $d" >> gpurun_out/lp_ew2.log; VQCPC_LP_DEBUG=$d python tools/lstm_time.py 2048 2>&1 | tail -2 >> gpurun_out/lp_ew2.log; done
