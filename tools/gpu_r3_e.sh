#!/bin/bash
# round 2, session 2, call e: K2 with shared zero flags + folded diagonal blocks (parity + timing), region build with 4 CTAs per SM
mkdir -p gpurun_out/r3
timeout 900 python -m pytest tests/test_gpu_pair_iou.py tests/test_gpu_parity_holes.py tests/test_gpu_parts_fullsize.py tests/test_gpu_jf_fused.py tests/test_gpu_abi_direct.py -x -q > gpurun_out/r3/pytest_e.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3/pytest_e.log
timeout 300 python tools/k2_bench.py > gpurun_out/r3/k2_bench_fold.json 2> gpurun_out/r3/k2_bench.err; cat gpurun_out/r3/k2_bench_fold.json
timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 100 > gpurun_out/r3/bench_e.json 2> gpurun_out/r3/bench_e.err; echo "rc=$?"; tail -2 gpurun_out/r3/bench_e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_e.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'roofline', round(d['roofline']['frac'],4), [round(v,3) for v in d['stage_ms'].values() if isinstance(v,float)], d['clocks']['sm_mhz'])
print('jf', d['roofline_jf']['frac'], d['roofline_jf']['ms_per_launch'], 'jfb', d['roofline_jf_boundary']['ms_per_launch'], 'k2', d['roofline_k2'])
PY
