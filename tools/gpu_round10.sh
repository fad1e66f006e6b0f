#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "conv2d_small or fpn4 or decoder" > gpurun_out/pytest_small.log 2>&1; echo "small/decoder tests rc=$?"
tail -3 gpurun_out/pytest_small.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_step.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_step.log 2>&1; echo "step tests rc=$?"; tail -3 gpurun_out/pytest_step.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline > gpurun_out/bench_r10.log 2> gpurun_out/bench_r10.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_r10.log | cut -c1-200
timeout 300 python tools/profile_step.py --ops > gpurun_out/step_profile.log 2>&1; grep -i "small_\|alg0\|grouped_direct\|cutlass__5x\|sm80_xmma" gpurun_out/step_profile.log | head -12 | cut -c1-150
