#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_ncu.log 2>&1; echo "ncu_rc=$?" >> gpurun_out/smoke_ncu.log
timeout 600 python -m pytest tests -m gpu -x -q -k "vocoder or smoke or generate" > gpurun_out/pytest_s4n.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_s4n.log
