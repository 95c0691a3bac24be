# round 2 evidence: launch list of the bench command, one full ncu capture of the step kernel, steady-state DRAM traffic,
# launch list of one PPO optimiser step, aux kernels.  Everything lands in gpurun_out/; summaries are copied to profiles/.
mkdir -p gpurun_out
export TA_COMMIT=$(cat gpurun_out/.commit 2>/dev/null || echo unknown)
timeout 300 python bench.py --steps 64 --warmup 8 --no-extra --no-cpu-baseline --no-ppo --e2e-steps 2 > gpurun_out/r2_plain_launches.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_v17.csv python bench.py --steps 64 --warmup 8 --no-extra --no-cpu-baseline --no-ppo --e2e-steps 2 > gpurun_out/r2_ncu_launches.log 2>&1
timeout 200 python scripts/prof_step.py > gpurun_out/r2_plain_step.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_obs -s 16 -c 2 -f -o gpurun_out/r2_step_obs_v17 python scripts/prof_step.py > gpurun_out/r2_ncu_step.log 2>&1
V=7 timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_obs -s 16 -c 2 -f -o gpurun_out/r2_step_obs_v7 python scripts/prof_step.py > gpurun_out/r2_ncu_step7.log 2>&1
K=40 timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:step_obs -s 16 -c 16 --csv --log-file gpurun_out/r2_step_obs_v17_traffic_steady.csv python scripts/prof_step.py > gpurun_out/r2_ncu_traffic.log 2>&1
K=40 V=7 timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:step_obs -s 16 -c 16 --csv --log-file gpurun_out/r2_step_obs_v7_traffic_steady.csv python scripts/prof_step.py > gpurun_out/r2_ncu_traffic7.log 2>&1
timeout 200 python scripts/prof_ppo_step.py > gpurun_out/r2_plain_ppo_step.log 2>&1 && \
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file gpurun_out/r2_ppo_step_ncu.csv python scripts/prof_ppo_step.py > gpurun_out/r2_ncu_ppo_step.log 2>&1
timeout 300 python bench.py --workload aux > gpurun_out/r2_aux_kernels.json 2> gpurun_out/r2_aux.err
ls -la gpurun_out | tail -20
