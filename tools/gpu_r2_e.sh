#!/bin/bash
mkdir -p gpurun_out/r2
cap() { local name=$1 rx=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o gpurun_out/r2/$name -f "$@" > gpurun_out/r2/$name.log 2>&1; echo "$name rc=$?"; }
timeout 600 python -m pytest tests/test_dataloader_adapter.py tests/test_rle.py tests/test_gpu_pair_iou.py tests/test_gpu_jf_fused.py -m gpu -x -q 2>&1 | tail -4
cap fused_bf16_720p fused_pack_resize_kernel 2 python tools/ncu_targets.py fused_bf16
cap fused_f32_480x854 band_pack_generic 2 python tools/ncu_targets.py fused_480p
cap fused_f32_480x864 fused_pack_resize_kernel 2 python tools/ncu_targets.py fused_480x864
cap k2_dense pair_iou_st_ring 2 python tools/ncu_targets.py k2_dense
