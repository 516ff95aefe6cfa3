#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "text_dump or encode_from_host" > gpurun_out/pytest_s4j.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_s4j.log
python tools/textdump_time.py > gpurun_out/textdump_time.log 2>&1
