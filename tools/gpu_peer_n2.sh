#!/bin/bash
# N-GPU check of the fused exchange + K2 (peer loads over NVLink) against the NCCL variants.   usage: gpu_peer_n2.sh N TRACKS [FRAMES]
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pair_iou.py -m gpu -q -k "row_pointer or parts" > gpurun_out/pytest_rows.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_rows.log; tail -3 gpurun_out/pytest_rows.log
N=${1:-2}; TR=${2:-64}; FR=${3:-200}
port=29533
IFS=";" read -ra MODE_LIST <<< "${MODES:---split words;--peer --chunks 4;--peer --chunks 2}"
for mode in "${MODE_LIST[@]}"; do
  tag=$(echo $mode | tr -d ' -')
  log=gpurun_out/stress_cfg5_n${N}_${tag}.log
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port tools/stress_cfg5_multigpu.py --tracks $TR --frames $FR --verify $mode > $log 2>&1; echo "rc=$?" >> $log
  grep -E '^\{|rc=|Error|error' $log | cut -c1-800
  port=$((port+1))
done
