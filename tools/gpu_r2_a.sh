#!/bin/bash
# round 2, call A: the GPU suite, the fused J&F kernel's bench + sanitizer + ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/jf_fused_bench.py > gpurun_out/r2_jf_fused_bench.json 2> gpurun_out/r2_jf_fused_bench.err; echo "jf bench rc=$?"; head -c 1500 gpurun_out/r2_jf_fused_bench.json
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_jf_fused.py -x -q > gpurun_out/r2_sanitizer_jf_fused.log 2>&1; echo "sanitizer rc=$?"; tail -4 gpurun_out/r2_sanitizer_jf_fused.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jf_fused_kernel -s 2 -c 1 -o gpurun_out/r2_jf_fused_720p -f python tools/jf_fused_ncu_target.py 720 1280 1280 > gpurun_out/ncu_jf.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jf_fused_kernel -s 2 -c 1 -o gpurun_out/r2_jf_fused_480p -f python tools/jf_fused_ncu_target.py 480 854 2048 > gpurun_out/ncu_jf2.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/ | head -30
