#!/bin/bash
# round 2, 8-GPU call: weak-scaling bench line at N = 8 and N = 4 (driver-style launch), e2e A/B with write-combined staging, host bandwidth
mkdir -p gpurun_out
nvidia-smi -L | wc -l
HB_GB=8 timeout 600 python tools/host_bandwidth.py > gpurun_out/r2_host_bandwidth_8gpu_box.json 2> gpurun_out/hb.err; cat gpurun_out/r2_host_bandwidth_8gpu_box.json
run() { local n=$1 tag=$2; shift 2
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus $n --steps 20 --warmup 3 "$@" > gpurun_out/r2_bench_${tag}.json 2> gpurun_out/r2_bench_${tag}.err
  echo "bench $tag rc=$?"; python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
for ln in open(f'gpurun_out/r2_bench_{tag}.json'):
    if ln.startswith('{'):
        d = json.loads(ln)
        e = d.get('e2e') or {}
        print(tag, 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 3), 'e2e', e.get('value') and round(e['value']), e.get('h2d_GBps_per_rank'), e.get('host_staging'))
        for k in ('jf_stage', 'cfg5'):
            if d.get(k): print('  ', k, json.dumps(d[k])[:700])
PY
}
run 8 8gpu
run 8 8gpu_wc --e2e-wc --no-jf --no-cfg5
run 4 4gpu --no-jf --no-cfg5
run 2 2gpu --no-jf --no-cfg5
