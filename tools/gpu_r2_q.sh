#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_gpu.log; tail -3 gpurun_out/r2/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2/smoke.log
