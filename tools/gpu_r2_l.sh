#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_jf_fused.py tests/test_gpu_boundary.py tests/test_gpu_abi_direct.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python tools/jf_fused_bench.py > gpurun_out/r2_jf_fused_bench_v6.json 2> gpurun_out/r2_jf_fused_bench.err; echo "jf bench rc=$?"; tail -2 gpurun_out/r2_jf_fused_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_jf_fused_bench_v6.json'))
for k,v in d.items():
    if 'boundary' in k: print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
PY
