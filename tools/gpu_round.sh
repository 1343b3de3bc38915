#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "pair or dedup or parts or fullsize" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/pytest_gpu.log | head -40
timeout 300 python tools/stage_breakdown.py > gpurun_out/stages.json 2>&1; grep -E "K2|pipelined|K1R1" gpurun_out/stages.json
timeout 900 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log; tail -2 gpurun_out/bench.log | cut -c1-220
