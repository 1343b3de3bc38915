#!/bin/bash
# round 2, session 2, call f: K2 with 16 consumer warps and a 4 x 2 micro-tile (parity + timing)
mkdir -p gpurun_out/r3
timeout 900 python -m pytest tests/test_gpu_pair_iou.py tests/test_gpu_parity_holes.py tests/test_gpu_parts_fullsize.py tests/test_gpu_abi_direct.py -x -q > gpurun_out/r3/pytest_f.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3/pytest_f.log
timeout 300 python tools/k2_bench.py > gpurun_out/r3/k2_bench_16warps_diag5.json 2> gpurun_out/r3/k2_bench.err; cat gpurun_out/r3/k2_bench_16warps_diag5.json
timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-jf --steps 100 > gpurun_out/r3/bench_f2.json 2> gpurun_out/r3/bench_f2.err; echo "rc=$?"; tail -2 gpurun_out/r3/bench_f2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_f2.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'roofline', round(d['roofline']['frac'],4), [round(v,3) for v in d['stage_ms'].values() if isinstance(v,float)], d['clocks']['sm_mhz'])
PY
