#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/pytest_gpu.log | head -40
timeout 300 python tools/stage_breakdown.py > gpurun_out/stages.json 2>&1; cat gpurun_out/stages.json
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log; tail -3 gpurun_out/bench.log
timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pack_flat_kernel -s 3 -c 1 -o gpurun_out/k1_full -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_k1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:resize_bilinear_tiled -s 3 -c 1 -o gpurun_out/r1_full -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_iou_st_kernel -s 3 -c 1 -o gpurun_out/k2_full -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_k2.log 2>&1
ls -la gpurun_out
