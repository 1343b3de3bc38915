#!/bin/bash
# round 2, session 2, call l: carry-save region counts (parity), re-capture of the kernels whose source changed after the r3 captures
mkdir -p gpurun_out/r3ncu gpurun_out/r3l
timeout 900 python -m pytest tests/test_gpu_jf_fused.py tests/test_gpu_counts.py tests/test_gpu_boundary.py tests/test_gpu_abi_direct.py tests/test_gpu_pair_iou.py -x -q > gpurun_out/r3l/pytest_l.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3l/pytest_l.log
timeout 300 python tools/jf_region_bench.py > gpurun_out/r3l/jf_region_bench.json 2>/dev/null; cat gpurun_out/r3l/jf_region_bench.json
cap() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/r3ncu/$name -f "$@" > gpurun_out/r3ncu/$name.log 2>&1
  echo "$name rc=$?"
}
cap gather pair_iou_gather 3 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-jf
cap jf_region jf_fused_kernel 2 python tools/ncu_targets.py jf_region
cap jf_boundary jf_fused_kernel 2 python tools/ncu_targets.py jf_boundary
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r3l/bench_1gpu.json 2> gpurun_out/r3l/bench_1gpu.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3l/bench_1gpu.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'roofline', round(d['roofline']['frac'],4), 'jf', round(d['roofline_jf']['frac'],4), d['roofline_jf']['ms_per_launch'], 'jfb', d['roofline_jf_boundary']['ms_per_launch'])
PY
