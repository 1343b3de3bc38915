#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "resize or dedup or fused" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/pytest_gpu.log | head -40
for wv in 2 4 8 16; do
SOLA_FUSED_WAVES=$wv timeout 900 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/bench_w$wv.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_w$wv.log") if x.startswith("{")]
d=json.loads(l[-1]); print("waves=$wv", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "fused ms", round(d["roofline"]["ms_per_launch"],3), "frac", round(d["roofline"]["frac"],3), d["clocks"])
PY
done
