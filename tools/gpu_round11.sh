#!/bin/bash
# ncu (a handful of metrics, CSV) of the own kernels inside one reg3d forward+backward: prob stencils, tcgen05 conv0 kernels,
# BatchNorm kernels (full-resolution three-kernel form and the one-kernel form of the coarser levels).
mkdir -p gpurun_out
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size \
    -k regex:"c16o1_fwd|c16o1_dgrad|c16o1_wgrad_kernel|c16c16_tc_kernel|c16c16_wgrad_tc|bn_fwd_fused|bn_bwd_fused|bn_apply_kernel|bn_bwd_reduce|bn_bwd_apply|bn_stats" \
    --csv --log-file gpurun_out/reg3d_own_kernels.csv python tools/ncu_reg3d_kernels.py 1 > gpurun_out/ncu_reg3d_own.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/reg3d_own_kernels.csv; ls -la gpurun_out/
