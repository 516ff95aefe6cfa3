#!/bin/bash
# repeated parity runs of the persistent LSTM (looking for a rare ordering bug in the release / acquire chain)
mkdir -p gpurun_out
: > gpurun_out/lp_stress.log
for i in 1 2 3 4 5 6; do
  python -m pytest tests -m gpu -x -q -k "lstm" 2>&1 | tail -1 >> gpurun_out/lp_stress.log
  python tools/lstm_check.py 64 100 300 512 1100 2304 4096 5000 2>&1 | awk '{print $1, $2, $NF}' | tr '\n' ' ' >> gpurun_out/lp_stress.log; echo >> gpurun_out/lp_stress.log
done
