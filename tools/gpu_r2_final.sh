#!/bin/bash
# round 2, final 1-GPU call: the full GPU suite, smoke(), both bench arms as the driver runs them, sanitizers, then the ncu evidence
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_gpu.log; tail -4 gpurun_out/r2/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2/smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2/bench_reference_arm.json 2> gpurun_out/r2/bench_reference_arm.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2/bench_reference_arm.json
timeout 900 python bench.py > gpurun_out/r2/bench_1gpu.json 2> gpurun_out/r2/bench_1gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r2/bench_1gpu.err; cut -c1-400 gpurun_out/r2/bench_1gpu.json
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/r2/compute_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r2/compute_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/r2/compute_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2/compute_sanitizer_racecheck.log
bash tools/gpu_profile_r2.sh
