#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "pair or dedup or parts or fullsize or survey" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/pytest_gpu.log | head
timeout 300 python tools/stage_breakdown.py > gpurun_out/stages.json 2>&1; grep -E "K2|pipelined" gpurun_out/stages.json
timeout 600 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench.log 2>&1; grep -m1 '^{' gpurun_out/bench.log | cut -c1-200
