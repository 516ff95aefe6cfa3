#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "vq" > gpurun_out/pytest_s4e.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_s4e.log
python tools/vq_flags.py 1000000 init > gpurun_out/vq_flags.log 2>&1
python tools/vq_flags.py 1000000 trained >> gpurun_out/vq_flags.log 2>&1
python tools/vq_exactness.py 4000000 >> gpurun_out/vq_flags.log 2>&1
(nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv -lms 100 > gpurun_out/smi.log &) 
python tools/vq_flags.py 8000000 trained >> gpurun_out/vq_flags.log 2>&1
python bench.py > gpurun_out/bench_s4e.json 2> gpurun_out/bench_s4e.err; echo "bench exit $?" >> gpurun_out/bench_s4e.err
