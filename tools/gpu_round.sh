#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "pair or dedup or parts or fullsize or survey" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=|trap|illegal" gpurun_out/pytest_gpu.log | head -20
timeout 200 python tools/stage_breakdown.py > gpurun_out/stages.json 2>&1; grep -E "K2|pipelined" gpurun_out/stages.json
SOLA_NO_TMA=1 timeout 200 python tools/stage_breakdown.py > gpurun_out/stages_notma.json 2>&1; grep -E "K2|pipelined" gpurun_out/stages_notma.json
