#!/bin/bash
# round 2, session 2 (2 GPUs): multi-rank tests (the peer-TMA ring with the 16-warp K2 layout), config-5 exchange variants, bench.py --gpus 2 both arms
mkdir -p gpurun_out/r3n2
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/r3n2/pytest_multirank.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3n2/pytest_multirank.log; tail -5 gpurun_out/r3n2/pytest_multirank.log
for mode in tma pull direct; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/stress_cfg5_multigpu.py --tracks 64 --frames 200 --peer --peer-mode $mode --verify > gpurun_out/r3n2/cfg5_2gpu_$mode.log 2>&1; echo "cfg5 $mode rc=$?"; grep -h '^{' gpurun_out/r3n2/cfg5_2gpu_$mode.log | cut -c1-700
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --impl reference --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3n2/bench_reference_2gpu.json 2> gpurun_out/r3n2/bench_reference_2gpu.err; echo "ref2 rc=$?"; cut -c1-200 gpurun_out/r3n2/bench_reference_2gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3n2/bench_2gpu.json 2> gpurun_out/r3n2/bench_2gpu.err; echo "bench2 rc=$?"; tail -3 gpurun_out/r3n2/bench_2gpu.err; python - <<'PY'
import json
for ln in open('gpurun_out/r3n2/bench_2gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln)
        for k in ('value','ms_per_step','e2e','roofline_jf','cfg5'): print(k, json.dumps(d.get(k))[:700])
PY
