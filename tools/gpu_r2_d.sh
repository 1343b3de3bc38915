#!/bin/bash
# round 2, call D (2 GPUs): multi-rank tests, peer-TMA kernel validation, bench.py --gpus 2 as the driver launches it
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_d.log; tail -15 gpurun_out/pytest_gpu_d.log
timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/multirank_worker.py --tma > gpurun_out/r2_peer_tma_validation.log 2>&1; echo "tma rc=$?"; grep MULTIRANK_RESULT gpurun_out/r2_peer_tma_validation.log; tail -3 gpurun_out/r2_peer_tma_validation.log
for mode in pull tma direct; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/stress_cfg5_multigpu.py --tracks 64 --frames 200 --peer --peer-mode $mode --verify > gpurun_out/r2_cfg5_2gpu_$mode.log 2>&1; echo "cfg5 $mode rc=$?"; grep -h '^{' gpurun_out/r2_cfg5_2gpu_$mode.log | cut -c1-700
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r2_bench_2gpu.err; python - <<'PY'
import json
for ln in open('gpurun_out/r2_bench_2gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln)
        for k in ('value','ms_per_step','e2e','jf_stage','roofline_jf','roofline_jf_boundary','cfg5'): print(k, json.dumps(d.get(k))[:900])
PY
