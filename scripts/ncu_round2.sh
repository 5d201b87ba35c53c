#!/bin/bash
# Round-2 profiler evidence (run under gpurun on ONE B200; numbers printed by runs under ncu are never bench values).
#   bash scripts/ncu_round2.sh            -> gpurun_out/r2_*.csv / .ncu-rep ; summarise here with scripts/ncu_traffic.py
mkdir -p gpurun_out
K='regex:prologue_kernel|linear_umma_kernel|gather_extra_kernel|attn_dense_kernel|attn_csr_vrow32_kernel|head_final'
# 1. launch lists (cold-cache, serialised: compare SHARES): the default workload (32 graphs) and the per-GPU share of the
#    8-GPU strong-scaling run (4 graphs = 3 600 nodes)
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip 60 -c 72 --csv --log-file gpurun_out/r2_launches_b32.csv \
    python bench.py --steps 6 --warmup 3 --e2e-loops 0 --no-cpu-baseline > gpurun_out/r2_ncu_b32.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip 60 -c 72 --csv --log-file gpurun_out/r2_launches_b4.csv \
    python bench.py --steps 6 --warmup 3 --e2e-loops 0 --no-cpu-baseline --global-batch 4 --no-auto-graph > gpurun_out/r2_ncu_b4.log 2>&1
# 2. full capture of two steps' kernels of the default workload
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 60 -c 38 -o gpurun_out/r2_full \
    python bench.py --steps 6 --warmup 3 --e2e-loops 0 --no-cpu-baseline > gpurun_out/r2_ncu_full.log 2>&1
ncu -i gpurun_out/r2_full.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,sm__maximum_warps_per_active_cycle_pct > gpurun_out/r2_full_summary.csv 2>&1
# everything the analysis needs as csv (all raw metrics; per-instruction stall samples of the two dense attention kernels),
# then drop the 80 MB report: gpurun only copies back 64 MiB
ncu -i gpurun_out/r2_full.ncu-rep --page raw --csv > gpurun_out/r2_full_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_full.ncu-rep --page source --csv --kernel-name regex:attn_dense_kernel --launch-count 1 > gpurun_out/r2_source_attn_dense_hidden.csv 2>/dev/null
rm -f gpurun_out/r2_full.ncu-rep
ls -la gpurun_out | tail -12
