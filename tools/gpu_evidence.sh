#!/bin/bash
# Evidence visit: ncu launch list of the bench command, ncu --set full of the cost-volume kernels (dram traffic), bench lines
# of BASELINE configs 3 and 5, reg3d timeline.
mkdir -p gpurun_out
timeout 1800 env MVD_REPORT=1 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_datapipe.py > gpurun_out/datapipe.log 2>&1; cat gpurun_out/datapipe.log
timeout 300 python tools/profile_step.py --ops > gpurun_out/step_profile.log 2>&1; head -5 gpurun_out/step_profile.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no_cpu_baseline --ncu_range > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:costvol_grouped -s 3 -c 2 -f -o gpurun_out/costvol_full \
    python tools/profile_costvol.py --iters 3 --bwd --layout 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 300 python tools/profile_costvol.py --bwd --layout 1 > gpurun_out/costvol_time.log 2>&1
timeout 300 python tools/profile_costvol.py --bwd --layout 1 --pose stress >> gpurun_out/costvol_time.log 2>&1; cat gpurun_out/costvol_time.log
timeout 900 python bench.py --config 3 --steps 5 --warmup 3 --no_cpu_baseline > gpurun_out/bench_c3.log 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"; tail -1 gpurun_out/bench_c3.log | cut -c1-200
timeout 900 python bench.py --config 5 --steps 5 --warmup 3 --no_cpu_baseline > gpurun_out/bench_c5.log 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"; tail -1 gpurun_out/bench_c5.log | cut -c1-200
timeout 300 python tools/profile_reg3d.py > gpurun_out/reg3d_timeline.log 2>&1; head -3 gpurun_out/reg3d_timeline.log
