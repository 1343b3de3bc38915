#!/bin/bash
# round 2, session 2, call g: where R2 runs (beside K1+R1 or beside K2), three repeats each
mkdir -p gpurun_out/r3
run() { tag=$1; shift
  timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-jf --steps 100 "$@" > gpurun_out/r3/bench_$tag.json 2> gpurun_out/r3/bench_$tag.err; echo "rc=$?"; tail -2 gpurun_out/r3/bench_$tag.err
  python - "$tag" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r3/bench_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'roofline', round(d['roofline']['frac'],4), [round(v,3) for v in d['stage_ms'].values() if isinstance(v,float)], d['clocks']['sm_mhz'])
PY
}
for i in 1 2; do
run g_early$i
run g_late$i --r2-late
run g_noaux$i --no-aux
done
