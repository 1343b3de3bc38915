#!/bin/bash
# round 2, session 2, last 1-GPU call on the final tree: the driver's sequence (GPU tests, smoke, reference arm, own arm)
mkdir -p gpurun_out/r3z
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3z/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3z/pytest_gpu.log; tail -4 gpurun_out/r3z/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3z/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r3z/smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/r3z/bench_reference_arm.json 2> gpurun_out/r3z/bench_reference_arm.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r3z/bench_reference_arm.json
timeout 900 python bench.py --gpus 1 > gpurun_out/r3z/bench_1gpu.json 2> gpurun_out/r3z/bench_1gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r3z/bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3z/bench_1gpu.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'roofline', round(d['roofline']['frac'],4), d['roofline']['traffic'], 'jf', round(d['roofline_jf']['frac'],4), 'jfb ms', d['roofline_jf_boundary']['ms_per_launch'], 'cpu', round(d['cpu_baseline']['value']), d['clocks'])
print(d['stage_ms'])
PY
timeout 300 python tools/k2_bench.py > gpurun_out/r3z/k2_bench.json 2>/dev/null; cat gpurun_out/r3z/k2_bench.json
