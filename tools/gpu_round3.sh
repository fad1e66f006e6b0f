#!/bin/bash
# GPU-box visit: parity tests, smoke, bench, step profile, tcgen05 conv time breakdown + ncu source capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
timeout 1800 env MVD_REPORT=1 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-400
timeout 300 python tools/profile_step.py --ops > gpurun_out/step_profile.log 2>&1; head -12 gpurun_out/step_profile.log
timeout 300 python tools/test_tc_conv.py --breakdown > gpurun_out/tc_breakdown.log 2>&1; tail -16 gpurun_out/tc_breakdown.log

timeout 300 python tools/bench_inference.py > gpurun_out/inference.log 2>&1; tail -1 gpurun_out/inference.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:c16c16_tc_kernel -s 12 -c 2 -f -o gpurun_out/tc_src \
    python tools/test_tc_conv.py > gpurun_out/ncu_tc_src.log 2>&1; echo "ncu rc=$?"
kill $SMI
