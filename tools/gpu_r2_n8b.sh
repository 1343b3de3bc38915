#!/bin/bash
# round 2, second 8-GPU call: FULL BASELINE config 5 (256 tracks x 200 frames x 1080x1920) through the three exchange paths, config-4 sweep, topology
mkdir -p gpurun_out
{ nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; free -g | head -2; } > gpurun_out/r2_topo_8gpu_box.txt 2>&1
for mode in "--peer --peer-mode tma" "--peer --peer-mode pull" "--split words"; do
  tag=$(echo $mode | tr -d ' -')
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29630 tools/stress_cfg5_multigpu.py --tracks 256 --frames 200 $mode --verify > gpurun_out/r2_cfg5_full_8gpu_$tag.log 2>&1; echo "cfg5 $tag rc=$?"; grep -h '^{' gpurun_out/r2_cfg5_full_8gpu_$tag.log | cut -c1-800
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 tools/jf_sweep_multigpu.py > gpurun_out/r2_jf_sweep_8gpu.log 2>&1; echo "jf sweep rc=$?"; grep -h '^{' gpurun_out/r2_jf_sweep_8gpu.log | cut -c1-700
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q 2>&1 | tail -2
head -14 gpurun_out/r2_topo_8gpu_box.txt
