#!/bin/bash
# round 2, session 2, call o: vectorised host formulas of the J&F sweep (GPU tests through JFSweep / Evaluator / adapter), bench second region
mkdir -p gpurun_out/r3o
timeout 900 python -m pytest tests/test_gpu_jf_fused.py tests/test_gpu_counts.py tests/test_dataloader_adapter.py tests/test_gpu_parts_fullsize.py -m gpu -x -q > gpurun_out/r3o/pytest_o.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r3o/pytest_o.log
timeout 600 python bench.py --no-e2e --no-cpu-baseline --steps 50 > gpurun_out/r3o/bench_o.json 2> gpurun_out/r3o/bench_o.err; echo "rc=$?"; tail -2 gpurun_out/r3o/bench_o.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3o/bench_o.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4))
for k,v in d['jf_stage'].items():
    if isinstance(v,dict): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('masklet_frames_per_s','kernel_only_masklet_frames_per_s','ms_per_sweep','kernel_ms','mean_J','mean_F')})
PY
