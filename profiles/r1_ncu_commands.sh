mkdir -p gpurun_out
# launch list (cold-cache, serialised): per-launch durations of one bench step
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 3 --warmup 3 --e2e-loops 0 --no-cpu-baseline > gpurun_out/r1b_ncu_bench.log 2>&1
# full capture of one step's kernels (skip the warm-up + first steps: 19 launches per step)
ncu --set full --clock-control none --import-source on --launch-skip 60 -c 20 -o gpurun_out/r1b_full python bench.py --steps 3 --warmup 3 --e2e-loops 0 --no-cpu-baseline > gpurun_out/r1b_ncu_full.log 2>&1
ncu -i gpurun_out/r1b_full.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum > gpurun_out/r1b_full_summary.csv 2>&1
ls -la gpurun_out | tail -8
