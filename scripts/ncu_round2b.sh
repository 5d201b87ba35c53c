#!/bin/bash
# Round-2 (second half) profiler evidence after the folded path / persistent hidden-layer kernel / table prologue.
# Run under gpurun on ONE B200; numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
K='regex:prologue|linear_umma_kernel|gather_extra_kernel|attn_dense|attn_hidden_persist|attn_fold|attn_csr_vrow32_kernel|head_f'
# 1. launch list (cold-cache, serialised: compare SHARES) of the default workload (32 graphs x 900 nodes)
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip 45 -c 60 --csv --log-file gpurun_out/r2b_launches_b32.csv \
    python bench.py --steps 6 --warmup 3 --e2e-loops 0 --no-cpu-baseline > gpurun_out/r2b_ncu_b32.log 2>&1
# 2. full capture of two steps' kernels
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 45 -c 30 -o gpurun_out/r2b_full \
    python bench.py --steps 6 --warmup 3 --e2e-loops 0 --no-cpu-baseline > gpurun_out/r2b_ncu_full.log 2>&1
ncu -i gpurun_out/r2b_full.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem > gpurun_out/r2b_full_summary.csv 2>&1
ncu -i gpurun_out/r2b_full.ncu-rep --page source --csv --kernel-name regex:linear_umma_kernel --launch-skip 1 --launch-count 1 > gpurun_out/r2b_source_gemm_mid.csv 2>/dev/null
ncu -i gpurun_out/r2b_full.ncu-rep --page source --csv --kernel-name regex:attn_hidden_persist --launch-count 1 > gpurun_out/r2b_source_attn_hidden_persist.csv 2>/dev/null
ncu -i gpurun_out/r2b_full.ncu-rep --page source --csv --kernel-name regex:attn_fold_persist --launch-count 1 > gpurun_out/r2b_source_attn_fold_persist.csv 2>/dev/null
rm -f gpurun_out/r2b_full.ncu-rep
ls -la gpurun_out | grep r2b
