#!/bin/bash
# round 2, session 2, final 1-GPU call: full GPU suite, smoke(), both bench arms as the driver runs them, sanitizers, ncu evidence
mkdir -p gpurun_out/r3f
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3f/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3f/pytest_gpu.log; tail -4 gpurun_out/r3f/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3f/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r3f/smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r3f/bench_reference_arm.json 2> gpurun_out/r3f/bench_reference_arm.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r3f/bench_reference_arm.json
timeout 900 python bench.py > gpurun_out/r3f/bench_1gpu.json 2> gpurun_out/r3f/bench_1gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/r3f/bench_1gpu.err; cut -c1-400 gpurun_out/r3f/bench_1gpu.json
timeout 300 python tools/k2_bench.py > gpurun_out/r3f/k2_bench.json 2>/dev/null; cat gpurun_out/r3f/k2_bench.json
timeout 300 python tools/jf_region_bench.py > gpurun_out/r3f/jf_region_bench.json 2>/dev/null; cat gpurun_out/r3f/jf_region_bench.json
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/r3f/compute_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r3f/compute_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/r3f/compute_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r3f/compute_sanitizer_racecheck.log
bash tools/gpu_profile_r3.sh
