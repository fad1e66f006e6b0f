#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline > gpurun_out/bench_r12.log 2> gpurun_out/bench_r12.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_r12.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["gpu_launches"])'
timeout 300 python tools/profile_step.py --ops > gpurun_out/step_profile.log 2>&1; grep -i "GPU busy\|vectorized_elementwise\|at::native::elementwise\|pose_matrix" gpurun_out/step_profile.log | head -6 | cut -c1-150
