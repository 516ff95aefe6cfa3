#!/bin/bash
# round-2 state capture: tests, bench, trace, ncu --set full of the three hot kernels
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
python tools/cl_trace.py s3 200 1 1 > gpurun_out/cl_trace.log 2>&1
python tools/vq_profile.py 1000000 init > gpurun_out/vq_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vq_tc_kernel -s 3 -c 1 -o gpurun_out/r02_vq_tc python tools/vq_profile.py 1000000 init > gpurun_out/ncu_vq.log 2>&1
python tools/encode_profile.py 4096 bf16x3 1 > gpurun_out/enc_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_ln_pair -s 12 -c 2 -o gpurun_out/r02_pair python tools/encode_profile.py 4096 bf16x3 1 > gpurun_out/ncu_pair.log 2>&1
python tools/gen_profile.py 1 50 1 > gpurun_out/gen_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ar_cluster -s 1 -c 1 -o gpurun_out/r02_ar_cluster python tools/gen_profile.py 1 50 1 > gpurun_out/ncu_ar.log 2>&1
ls -la gpurun_out
