#!/bin/bash
# round 2, session 2, 8-GPU call: driver-style bench at N = 8 (own arm, e2e skipped: it is host-bound and unchanged, profiles/r2_bench_8gpu.json)
# and the FULL BASELINE config 5 through the one-kernel exchange + K2 with the new K2 layout
mkdir -p gpurun_out/r3n8
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e > gpurun_out/r3n8/bench_8gpu.json 2> gpurun_out/r3n8/bench_8gpu.err
echo "bench 8 rc=$?"; python - <<'PY'
import json
for ln in open('gpurun_out/r3n8/bench_8gpu.json'):
    if ln.startswith('{'):
        d = json.loads(ln)
        print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3))
        for k in ('jf_stage', 'cfg5'):
            if d.get(k): print('  ', k, json.dumps(d[k])[:900])
PY
for mode in "--peer --peer-mode tma" "--split words"; do
  tag=$(echo $mode | tr -d ' -')
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29630 tools/stress_cfg5_multigpu.py --tracks 256 --frames 200 $mode --verify > gpurun_out/r3n8/cfg5_full_8gpu_$tag.log 2>&1; echo "cfg5 $tag rc=$?"; grep -h '^{' gpurun_out/r3n8/cfg5_full_8gpu_$tag.log | cut -c1-800
done
