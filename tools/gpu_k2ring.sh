#!/bin/bash
mkdir -p gpurun_out
SOLA_K2_RING=1 timeout 300 python -m pytest tests/test_gpu_pair_iou.py tests/test_gpu_parts_fullsize.py tests/test_gpu_abi_direct.py -m gpu -x -q > gpurun_out/pytest_ring.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ring.log; tail -3 gpurun_out/pytest_ring.log
for v in 0 1; do SOLA_K2_RING=$v timeout 120 python tools/k2_bench.py 2>&1 | tail -1 | sed "s/^{/{\"ring\": $v, /"; done | tee gpurun_out/k2_ring.jsonl
SOLA_K2_RING=1 timeout 300 python bench.py --no-e2e --no-cpu-baseline 2>&1 | grep -m1 '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ring bench', d['value'], d['ms_per_step'], d['stage_ms'])"
