#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  MVD_CUDNN_BENCHMARK=$v timeout 900 python bench.py --steps 10 --warmup 4 --no_cpu_baseline > gpurun_out/bench_cb$v.log 2> gpurun_out/bench_cb$v.err
  echo "cudnn.benchmark=$v rc=$? $(tail -1 gpurun_out/bench_cb$v.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"])' 2>&1 | tail -1)"
done
MVD_CUDNN_BENCHMARK=1 timeout 900 python -m pytest tests/test_gpu_step.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_step_cb.log 2>&1; echo "step tests (benchmark on) rc=$?"; tail -2 gpurun_out/pytest_step_cb.log | cut -c1-200
