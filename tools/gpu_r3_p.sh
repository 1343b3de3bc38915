#!/bin/bash
# round 2, session 2, call p: threads per CTA of the fused J&F kernel (256 / 384 / 512; two CTAs per SM): parity + per-shape throughput
mkdir -p gpurun_out/r3p
for n in 256 512 384; do
  export SOLA_EXTRA_NVCC_FLAGS="-DJF_THREADS_VALUE=$n"
  timeout 600 python -m pytest tests/test_gpu_jf_fused.py tests/test_gpu_boundary.py -x -q > gpurun_out/r3p/pytest_$n.log 2>&1; echo "threads $n pytest rc=$?"; tail -1 gpurun_out/r3p/pytest_$n.log
  timeout 600 python tools/jf_fused_bench.py --auto-only > gpurun_out/r3p/jf_fused_bench_$n.json 2> gpurun_out/r3p/jf_$n.err
  python - $n <<'PY'
import json,sys
d=json.load(open('gpurun_out/r3p/jf_fused_bench_%s.json'%sys.argv[1]))
print(sys.argv[1], {k: round(v['frames_per_s']/1e6,3) for k,v in d.items() if 'J+F+boundary' in k and 'frames_per_s' in v})
PY
done
unset SOLA_EXTRA_NVCC_FLAGS
python -c "import sola_b200._build as b; b.build()"
