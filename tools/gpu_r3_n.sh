#!/bin/bash
# round 2, session 2, call n: R2 rows per warp item (parity + timing), then the step
mkdir -p gpurun_out/r3n
timeout 600 python -m pytest tests/test_gpu_resize.py tests/test_gpu_pair_iou.py -x -q > gpurun_out/r3n/pytest_n.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3n/pytest_n.log
for n in 4 8 16 32; do
  SOLA_EXTRA_NVCC_FLAGS="-DNN_ROWS_VALUE=$n" timeout 300 python tools/r2_bench.py 2>gpurun_out/r3n/r2_$n.err | tee -a gpurun_out/r3n/r2_rows.jsonl
done
python -c "import sola_b200._build as b; b.build()"
timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-jf --steps 100 > gpurun_out/r3n/bench_n.json 2> gpurun_out/r3n/bench_n.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3n/bench_n.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'roofline', round(d['roofline']['frac'],4), [round(v,3) for v in d['stage_ms'].values() if isinstance(v,float)], d['clocks']['sm_mhz'])
PY
