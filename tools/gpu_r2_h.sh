#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_jf_fused.py tests/test_gpu_boundary.py tests/test_gpu_parts_fullsize.py tests/test_dataloader_adapter.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/jf_fused_bench.py > gpurun_out/r2_jf_fused_bench_v5.json 2> gpurun_out/r2_jf_fused_bench.err; echo "jf bench rc=$?"; tail -2 gpurun_out/r2_jf_fused_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_jf_fused_bench_v5.json'))
for k,v in d.items():
    if 'boundary' in k: print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jf_fused_kernel -s 2 -c 1 -o gpurun_out/r2_jf_fused_720p_v5 -f python tools/jf_fused_ncu_target.py 720 1280 1280 > gpurun_out/ncu_jf.log 2>&1; echo "ncu rc=$?"
