#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, kernel timings, ncu launch list + full capture of K1.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
timeout 300 python tools/profile_costvol.py --bwd > gpurun_out/costvol_time.log 2>&1; cat gpurun_out/costvol_time.log
timeout 300 python tools/profile_costvol.py --bwd --pose stress >> gpurun_out/costvol_time.log 2>&1
timeout 300 python tools/profile_step.py > gpurun_out/step_profile.log 2>&1; head -30 gpurun_out/step_profile.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no_cpu_baseline --ncu_range > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:costvol_grouped -s 3 -c 2 -f -o gpurun_out/costvol_full \
    python tools/profile_costvol.py --iters 3 --bwd > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
kill $SMI
