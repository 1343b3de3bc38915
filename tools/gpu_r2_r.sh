#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_gpu.log; tail -3 gpurun_out/r2/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2/bench_1gpu.json 2> gpurun_out/r2/bench_1gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r2/bench_1gpu.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_1gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'])
PY
