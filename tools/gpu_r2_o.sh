#!/bin/bash
# experiment: radius of the fused J&F kernel's pre-test (build-time constant)
mkdir -p gpurun_out
for R in 2 3 4; do
  export SOLA_EXTRA_NVCC_FLAGS="-DJF_PRE_R_VALUE=$R"
  python -m sola_b200._build --force > /dev/null 2>&1
  timeout 600 python tools/jf_fused_bench.py --auto-only > gpurun_out/r2_jf_pre_r$R.json 2> gpurun_out/r2_jf_pre.err; echo "pre_r=$R rc=$?"
  python - $R <<'PY'
import json, sys
d=json.load(open(f'gpurun_out/r2_jf_pre_r{sys.argv[1]}.json'))
print(' '.join(f"{k.split('/')[0]}/{k.split('/')[1][:3]}={v['frames_per_s']/1e6:.2f}M" for k,v in d.items() if k.endswith('J+F+boundary')))
PY
done
