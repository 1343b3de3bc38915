#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "resize or nearest or dedup" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/pytest_gpu.log | head
timeout 300 python tools/stage_breakdown.py > gpurun_out/stages.json 2>&1; grep -E "R2|pipelined" gpurun_out/stages.json
