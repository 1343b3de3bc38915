#!/bin/bash
# round 2, session 2, call d: region-only build of the fused J&F kernel (parity, CTAs-per-SM sweep)
mkdir -p gpurun_out/r3
timeout 900 python -m pytest tests/test_gpu_jf_fused.py tests/test_gpu_counts.py tests/test_gpu_boundary.py -x -q > gpurun_out/r3/pytest_jf.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3/pytest_jf.log
for n in 8 6 4 3 2; do
  SOLA_EXTRA_NVCC_FLAGS="-DJF_REGION_CTAS_VALUE=$n" timeout 300 python tools/jf_region_bench.py 2>gpurun_out/r3/jf_region_$n.err | tee -a gpurun_out/r3/jf_region_ctas.jsonl
done
python -c "import sola_b200._build as b; b.build()"   # back to the default build
