#!/bin/bash
# GPU-box visit: one-kernel BatchNorm -- its tests first (bounded), then the whole suite, then the bench A/B over the fuse threshold.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "batchnorm" > gpurun_out/pytest_bn.log 2>&1; echo "bn tests rc=$?"
tail -4 gpurun_out/pytest_bn.log
for mb in 0 8 24 64; do
  MVD_BN_FUSE_MB=$mb timeout 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline > gpurun_out/bench_fuse$mb.log 2> gpurun_out/bench_fuse$mb.err
  echo "fuse<=${mb}MB rc=$? $(tail -1 gpurun_out/bench_fuse$mb.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["gpu_launches"])' 2>&1 | tail -1)"
done
timeout 1800 env MVD_REPORT=1 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_reg3d.py > gpurun_out/reg3d_timeline.log 2>&1; head -3 gpurun_out/reg3d_timeline.log
