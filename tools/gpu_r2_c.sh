#!/bin/bash
# round 2, call C: fused J&F v3 (mask work lists), new bench.py (both arms)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_jf_fused.py tests/test_gpu_boundary.py tests/test_gpu_abi_direct.py tests/test_gpu_counts.py -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_c.log; tail -15 gpurun_out/pytest_gpu_c.log
timeout 600 python tools/jf_fused_bench.py > gpurun_out/r2_jf_fused_bench_v3.json 2> gpurun_out/r2_jf_fused_bench.err; echo "jf bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_jf_fused_bench_v3.json'))
for k,v in d.items():
    if 'boundary' in k: print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jf_fused_kernel -s 2 -c 1 -o gpurun_out/r2_jf_fused_720p_v3 -f python tools/jf_fused_ncu_target.py 720 1280 1280 > gpurun_out/ncu_jf.log 2>&1; echo "ncu rc=$?"
timeout 900 python bench.py --steps 30 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_c.err; head -c 3000 gpurun_out/r2_bench_c.json; echo
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_ref_c.json 2> gpurun_out/r2_bench_ref_c.err; echo "ref rc=$?"; tail -3 gpurun_out/r2_bench_ref_c.err; head -c 2500 gpurun_out/r2_bench_ref_c.json; echo
