#!/bin/bash
# GPU-box visit: K1b v5 -- race statistics, parity, timing against the round-1 kernel, ncu, bench.
mkdir -p gpurun_out
timeout 400 python tools/debug_k1b.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "costvol" > gpurun_out/pytest_costvol.log 2>&1; echo "costvol tests rc=$?"
tail -3 gpurun_out/pytest_costvol.log | cut -c1-200
for pose in forward stress; do
  for fl in 0 512 1024 32; do
    echo "pose=$pose flags=$fl $(timeout 300 python tools/profile_costvol.py --bwd --layout 1 --pose $pose --flags $fl 2>&1 | grep -i "bwd" | tail -1)"
  done
done
timeout 300 python tools/profile_costvol.py --bwd --layout 1 > gpurun_out/costvol_time.log 2>&1
timeout 300 python tools/profile_costvol.py --bwd --layout 1 --pose stress >> gpurun_out/costvol_time.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:costvol_grouped -s 3 -c 2 -f -o gpurun_out/costvol_full \
    python tools/profile_costvol.py --iters 3 --bwd --layout 1 > gpurun_out/ncu_costvol.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline > gpurun_out/bench_v5.log 2> gpurun_out/bench_v5.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_v5.log | cut -c1-200
