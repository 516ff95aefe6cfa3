#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "encode_from_host or lstm or encoder" > gpurun_out/pytest_s4i.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_s4i.log
python bench.py --workload encode_4096 --steps 5 --warmup 3 > gpurun_out/bench_enc4096.json 2> gpurun_out/bench_enc4096.err; echo "exit $?" >> gpurun_out/bench_enc4096.err
