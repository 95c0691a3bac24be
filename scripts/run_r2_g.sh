mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8
timeout 200 python scripts/prof_ppo_step.py > gpurun_out/r2_plain_ppo_step.log 2>&1 && \
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file gpurun_out/r2_ppo_step_ncu.csv python scripts/prof_ppo_step.py > gpurun_out/r2_ncu_ppo_step.log 2>&1
python scripts/ncu_launch_summary.py gpurun_out/r2_ppo_step_ncu.csv 45 > gpurun_out/r2_ppo_step_ncu_summary.txt; cat gpurun_out/r2_ppo_step_ncu_summary.txt
