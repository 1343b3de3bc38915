#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
for v in 0 2 3; do SOLA_K2_PLAIN=$v timeout 120 python tools/k2_bench.py 2>&1 | tail -1; done | tee gpurun_out/k2_bench.jsonl
timeout 300 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; grep -m1 '^{' gpurun_out/bench_quick.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['stage_ms'], d['roofline']['frac'])"
timeout 200 python tools/bf16_bench.py > gpurun_out/bf16_bench.json 2>&1; python -c "
import json; d=json.load(open('gpurun_out/bf16_bench.json')); print({k:(round(v['ms'],3), round(v.get('GBps',0))) for k,v in d.items()})"
timeout 200 python tools/generic_path_bench.py > gpurun_out/generic_bench.json 2>&1; python -c "
import json; d=json.load(open('gpurun_out/generic_bench.json')); print({k:(round(v['ms'],3), round(v.get('GBps',0))) for k,v in d.items()})"
