#!/bin/bash
# round 2, session 2, call b: two-stream step with K2 limited to one CTA per SM
mkdir -p gpurun_out/r3
run() { tag=$1; shift
  timeout 300 env "$@" python bench.py --no-e2e --no-cpu-baseline --no-jf --steps 100 $EXTRA > gpurun_out/r3/bench_$tag.json 2> gpurun_out/r3/bench_$tag.err; echo "rc=$?"; tail -2 gpurun_out/r3/bench_$tag.err
  python - "$tag" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r3/bench_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'roofline', round(d['roofline']['frac'],4), [round(v,3) for v in d['stage_ms'].values() if isinstance(v,float)], d['clocks']['sm_mhz'])
PY
}
EXTRA=""
run solo1_p-1 SOLA_K2_SOLO=1 SOLA_TAIL_PRIO=-1
run solo1_p0 SOLA_K2_SOLO=1 SOLA_TAIL_PRIO=0
run solo2_p-1 SOLA_K2_SOLO=2 SOLA_TAIL_PRIO=-1
run solo0_p0 SOLA_K2_SOLO=0 SOLA_TAIL_PRIO=0
EXTRA="--no-overlap"
run serial_solo1 SOLA_K2_SOLO=1
run serial_solo2 SOLA_K2_SOLO=2
