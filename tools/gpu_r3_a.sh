#!/bin/bash
# round 2, session 2, call a: two-stream step (tail of video k under K1+R1 of video k+1) against the serial step
mkdir -p gpurun_out/r3
for f in "" "--no-overlap"; do
  timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-jf --steps 100 $f > gpurun_out/r3/bench_overlap$f.json 2> gpurun_out/r3/bench_overlap$f.err; echo "rc=$?"; tail -2 gpurun_out/r3/bench_overlap$f.err
  python - "$f" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r3/bench_overlap%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1] or 'overlap', 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'roofline', round(d['roofline']['frac'],4), d['stage_ms'], d['clocks'])
PY
done
