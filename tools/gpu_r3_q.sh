#!/bin/bash
# round 2, session 2, call q: mbarrier.try_wait suspend-time hint (spinning waiters were 26 % of K2's warp instructions on object-like masklets)
mkdir -p gpurun_out/r3q
for n in 0 200 2000 20000; do
  SOLA_EXTRA_NVCC_FLAGS="-DMBAR_SUSPEND_NS=$n" timeout 300 python tools/k2_bench.py 2>gpurun_out/r3q/k2_$n.err | sed "s/^/{\"suspend_ns\": $n, \"r\": /; s/$/}/" | tee -a gpurun_out/r3q/k2_suspend.jsonl
done
python -c "import sola_b200._build as b; b.build()"
