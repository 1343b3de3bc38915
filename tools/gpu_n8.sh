#!/bin/bash
# 8-GPU round: host topology, config-5 exchange variants, the N=8 bench line (with and without GPU-local CPU binding).
mkdir -p gpurun_out
{ nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; python -c "import os; print('affinity', len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4], '...')"; free -g | head -2; } > gpurun_out/topo_n8.txt 2>&1
N=${1:-8}
MODES="--split words;--peer --chunks 4;--peer --chunks 2" bash tools/gpu_peer_n2.sh $N $((32*N)) 200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N > gpurun_out/bench_n${N}.log 2>&1; grep -m1 '^{' gpurun_out/bench_n${N}.log | cut -c1-2500
SOLA_BENCH_NO_AFFINITY=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 20 > gpurun_out/bench_n${N}_noaff.log 2>&1; grep -m1 '^{' gpurun_out/bench_n${N}_noaff.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('no-affinity', d['value'], d['e2e'])"
head -40 gpurun_out/topo_n8.txt
