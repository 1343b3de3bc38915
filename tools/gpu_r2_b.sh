#!/bin/bash
# round 2, call B: new tests, fused J&F v2 bench + ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_jf_fused.py tests/test_gpu_boundary.py tests/test_gpu_parity_holes.py tests/test_gpu_pair_iou.py tests/test_gpu_abi_direct.py -m gpu -x -q > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_b.log; tail -15 gpurun_out/pytest_gpu_b.log
timeout 600 python tools/jf_fused_bench.py > gpurun_out/r2_jf_fused_bench_v2.json 2> gpurun_out/r2_jf_fused_bench.err; echo "jf bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_jf_fused_bench_v2.json'))
for k,v in d.items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jf_fused_kernel -s 2 -c 1 -o gpurun_out/r2_jf_fused_720p_v2 -f python tools/jf_fused_ncu_target.py 720 1280 1280 > gpurun_out/ncu_jf.log 2>&1; echo "ncu rc=$?"
