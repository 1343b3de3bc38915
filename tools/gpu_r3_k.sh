#!/bin/bash
# round 2, session 2, call k: build-constant A/B — K2 ring depth, label track tile, carry-save region counts; re-capture of the gather kernel
mkdir -p gpurun_out/r3k
for n in 4 5 6; do
  SOLA_EXTRA_NVCC_FLAGS="-DK2_NSTAGE_VALUE=$n" timeout 300 python tools/k2_bench.py 2>gpurun_out/r3k/k2_$n.err | sed "s/^/{\"nstage\": $n, \"r\": /; s/$/}/" | tee -a gpurun_out/r3k/k2_nstage.jsonl
done
for f in "-DLABEL_NA_TILE=2" "-DLABEL_NA_TILE=4 -DLABEL_MIN_CTAS=2" "-DLABEL_NA_TILE=4 -DLABEL_MIN_CTAS=3"; do
  SOLA_EXTRA_NVCC_FLAGS="$f" timeout 300 python tools/labels_bench.py 2>gpurun_out/r3k/labels.err | tee -a gpurun_out/r3k/labels_na_tile.jsonl
done
for f in "" "-DJF_REGION_CSA"; do
  SOLA_EXTRA_NVCC_FLAGS="$f" timeout 300 python tools/jf_region_bench.py 2>gpurun_out/r3k/jfr.err | tee -a gpurun_out/r3k/jf_region_csa.jsonl
done
python -c "import sola_b200._build as b; b.build()"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pair_iou_gather -s 3 -c 1 -o gpurun_out/r3ncu/gather -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-jf > gpurun_out/r3ncu/gather.log 2>&1; echo "gather cap rc=$?"
