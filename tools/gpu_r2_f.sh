#!/bin/bash
# round 2, call F (1 GPU): fused K1+R1 on other shapes / dtypes, bench e2e with write-combined staging, host bandwidth
mkdir -p gpurun_out
timeout 900 python tools/fused_shapes_bench.py > gpurun_out/r2_fused_shapes_bench.json 2> gpurun_out/r2_fused_shapes.err; echo "shapes rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_fused_shapes_bench.json'))
for k,v in d.items(): print(f"{k:28s} K1 {v['K1_GBps']:7.0f} GB/s   fused {v['fused_GBps']:7.0f} GB/s = {v['fused_frac_of_6555.8']:.2f}")
PY
timeout 600 python bench.py --steps 20 --no-jf --no-cpu-baseline --e2e-wc > gpurun_out/r2_bench_e2e_wc.json 2> gpurun_out/r2_bench_e2e_wc.err; echo "wc rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_e2e_wc.json').read().strip().splitlines()[-1]); print('e2e wc', d['e2e'])"
HB_GB=4 timeout 300 python tools/host_bandwidth.py > gpurun_out/r2_host_bandwidth_1gpu_box.json; cat gpurun_out/r2_host_bandwidth_1gpu_box.json
