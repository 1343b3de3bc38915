#!/bin/bash
# round 2, session 2, call c: K2 strided stage walk (parity + timing), aux-stream step against the one-stream step
mkdir -p gpurun_out/r3
timeout 900 python -m pytest tests/test_gpu_pair_iou.py tests/test_gpu_parity_holes.py tests/test_gpu_parts_fullsize.py -x -q > gpurun_out/r3/pytest_k2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3/pytest_k2.log
timeout 300 python tools/k2_bench.py > gpurun_out/r3/k2_bench_strided.json 2> gpurun_out/r3/k2_bench.err; cat gpurun_out/r3/k2_bench_strided.json
run() { tag=$1; shift
  timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-jf --steps 100 "$@" > gpurun_out/r3/bench_$tag.json 2> gpurun_out/r3/bench_$tag.err; echo "rc=$?"; tail -2 gpurun_out/r3/bench_$tag.err
  python - "$tag" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r3/bench_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'roofline', round(d['roofline']['frac'],4), [round(v,3) for v in d['stage_ms'].values() if isinstance(v,float)], d['clocks']['sm_mhz'])
PY
}
run c_aux
run c_noaux --no-aux
run c_aux_overlap --overlap
