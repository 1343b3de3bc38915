#!/bin/bash
# round 2, session 2, call j: full GPU suite + smoke + the step with read-backs on the aux stream
mkdir -p gpurun_out/r3
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3/pytest_gpu.log; tail -4 gpurun_out/r3/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r3/smoke.log
for i in 1 2; do
timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-jf --steps 100 > gpurun_out/r3/bench_j$i.json 2> gpurun_out/r3/bench_j$i.err; echo "rc=$?"; tail -2 gpurun_out/r3/bench_j$i.err
python - $i <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r3/bench_j%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'roofline', round(d['roofline']['frac'],4), [round(v,3) for v in d['stage_ms'].values() if isinstance(v,float)], d['clocks']['sm_mhz'])
PY
done
