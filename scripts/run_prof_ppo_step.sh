# launch list of one PPO optimiser step with tensor-pipe / DRAM utilisation per kernel (ncu, no replay-heavy sets)
timeout 200 python scripts/prof_ppo_step.py > gpurun_out/plain_ppo_step.log 2>&1 && \
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file gpurun_out/ppo_step_ncu.csv python scripts/prof_ppo_step.py > gpurun_out/ncu_ppo_step.log 2>&1
tail -2 gpurun_out/ncu_ppo_step.log; wc -l gpurun_out/ppo_step_ncu.csv
