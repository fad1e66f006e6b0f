#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "conv3d or prob or c16 or reg3d or conv2d_small or fpn" > gpurun_out/pytest_conv.log 2>&1; echo "conv tests rc=$?"
tail -2 gpurun_out/pytest_conv.log | cut -c1-200
timeout 300 python tools/profile_reg3d.py > gpurun_out/reg3d_timeline.log 2>&1; grep -i "reg3d fwd\|c16o1" gpurun_out/reg3d_timeline.log | head
timeout 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline > gpurun_out/bench_r7.log 2> gpurun_out/bench_r7.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_r7.log | cut -c1-200
