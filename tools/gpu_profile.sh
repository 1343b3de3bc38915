#!/bin/bash
# ncu evidence for profiles/: GPU tests, bench line, launch list of the timed region, full captures of the dominant kernels.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log; grep -m1 '^{' gpurun_out/bench.log | cut -c1-400
timeout 300 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_pack_resize_kernel -s 3 -c 1 -o gpurun_out/fused_full -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_fused.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pair_iou_st_ -s 3 -c 1 -o gpurun_out/k2_full -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_k2.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
