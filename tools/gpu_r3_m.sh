#!/bin/bash
# round 2, session 2, call m: K2 grid size (CTAs launched per SM: waves of two resident CTAs)
mkdir -p gpurun_out/r3m
for n in 2 4 6 8; do
  SOLA_EXTRA_NVCC_FLAGS="-DK2_CTAS_PER_SM_TOTAL_VALUE=$n" timeout 300 python tools/k2_bench.py 2>gpurun_out/r3m/k2_$n.err | sed "s/^/{\"ctas_per_sm_total\": $n, \"r\": /; s/$/}/" | tee -a gpurun_out/r3m/k2_grid.jsonl
done
python -c "import sola_b200._build as b; b.build()"
