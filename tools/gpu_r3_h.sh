#!/bin/bash
# round 2, session 2, call h: carry-save label counts (parity + timing at 3 / 4 / 5 CTAs per SM), then the step
mkdir -p gpurun_out/r3
timeout 600 python -m pytest tests/test_gpu_counts.py tests/test_gpu_pair_iou.py -x -q > gpurun_out/r3/pytest_h.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3/pytest_h.log
for n in 3 5 4; do
  SOLA_EXTRA_NVCC_FLAGS="-DLABEL_MIN_CTAS=$n" timeout 300 python tools/labels_bench.py 2>gpurun_out/r3/labels_$n.err | tee -a gpurun_out/r3/labels_ctas.jsonl
done
python -c "import sola_b200._build as b; b.build()"
timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-jf --steps 100 > gpurun_out/r3/bench_h.json 2> gpurun_out/r3/bench_h.err; echo "rc=$?"; tail -2 gpurun_out/r3/bench_h.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3/bench_h.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'roofline', round(d['roofline']['frac'],4), [round(v,3) for v in d['stage_ms'].values() if isinstance(v,float)], d['clocks']['sm_mhz'])
PY
