#!/bin/bash
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_gpu.log; tail -3 gpurun_out/r2/pytest_gpu.log
timeout 900 python tools/jf_fused_bench.py > gpurun_out/r2_jf_fused_bench_v8.json 2> gpurun_out/r2_jf_fused_bench.err; echo "jf bench rc=$?"
cap() { local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/r2/$name -f "$@" > gpurun_out/r2/$name.log 2>&1; echo "$name rc=$?"; }
cap jf_region jf_fused_kernel 2 python tools/ncu_targets.py jf_region
cap jf_boundary jf_fused_kernel 2 python tools/ncu_targets.py jf_boundary
timeout 900 python bench.py > gpurun_out/r2/bench_1gpu.json 2> gpurun_out/r2/bench_1gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r2/bench_1gpu.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_1gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step']); print('e2e', d['e2e']['value']); print(json.dumps(d['jf_stage'])[:700]); print(d['roofline_jf']['frac'], d['roofline_jf_boundary']['frac'])
PY
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/r2/compute_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r2/compute_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/r2/compute_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r2/compute_sanitizer_racecheck.log
