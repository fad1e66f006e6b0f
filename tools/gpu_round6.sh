#!/bin/bash
# GPU-box visit: packed-FMA kernels -- whole suite, reg3d timeline, step profile, bench.
mkdir -p gpurun_out
timeout 1800 env MVD_REPORT=1 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 300 python tools/profile_reg3d.py > gpurun_out/reg3d_timeline.log 2>&1; grep -i "reg3d fwd\|c16o1" gpurun_out/reg3d_timeline.log | head
timeout 300 python tools/profile_step.py --ops > gpurun_out/step_profile.log 2>&1; grep -i "small_\|c16o1\|costvol" gpurun_out/step_profile.log | head -12 | cut -c1-150
timeout 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline > gpurun_out/bench_r6.log 2> gpurun_out/bench_r6.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_r6.log | cut -c1-200
