#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/pytest_gpu.log | head -40
timeout 300 python tools/stage_breakdown.py > gpurun_out/stages.json 2>&1; cat gpurun_out/stages.json
for fl in "" "--no-fuse"; do
timeout 900 python bench.py --no-e2e --no-cpu-baseline $fl > gpurun_out/bench$fl.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench$fl.log; tail -2 gpurun_out/bench$fl.log | cut -c1-220; grep -o '"roofline".*' gpurun_out/bench$fl.log | cut -c1-600
done
