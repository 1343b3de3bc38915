#!/bin/bash
# round-2 (second session) evidence for profiles/: launch list of the bench's timed region + `ncu --set full` captures of every kernel on the path.
mkdir -p gpurun_out/r3ncu
cap() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/r3ncu/$name -f "$@" > gpurun_out/r3ncu/$name.log 2>&1
  echo "$name rc=$?"
}
timeout 400 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3ncu/launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-jf > gpurun_out/r3ncu/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
cap fused_f32_720p fused_pack_resize_kernel 3 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-jf
cap k2_object pair_iou_st_ring 3 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-jf
cap labels packed_counts_kernel 3 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-jf
cap gather pair_iou_gather 3 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-jf
cap fused_bf16_720p fused_pack_resize_kernel 2 python tools/ncu_targets.py fused_bf16
cap fused_f32_480x854 band_pack_generic 2 python tools/ncu_targets.py fused_480p
cap fused_f32_480x864 fused_pack_resize_kernel 2 python tools/ncu_targets.py fused_480x864
cap k3_f32 raw_counts_vec 2 python tools/ncu_targets.py k3_f32
cap k2_dense pair_iou_st_ring 2 python tools/ncu_targets.py k2_dense
cap jf_region jf_fused_kernel 2 python tools/ncu_targets.py jf_region
cap jf_boundary jf_fused_kernel 2 python tools/ncu_targets.py jf_boundary
cap rle_fill rle_fill_runs 2 python tools/ncu_targets.py rle_decode
ls -la gpurun_out/r3ncu | head -40
cap resize_nearest resize_nearest_kernel 3 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-jf
