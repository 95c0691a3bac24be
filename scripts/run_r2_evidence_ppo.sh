# round-2 evidence at the final kernels: launch list of ONE optimiser step, --set full captures of the tcgen05 kernels,
# role profile of the fused stem-backward kernel, timelines
mkdir -p gpurun_out
timeout 300 python scripts/prof_ppo_step.py > gpurun_out/r2_plain_ppo_step.log 2>&1 && \
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file gpurun_out/r2_ppo_step_ncu.csv python scripts/prof_ppo_step.py > gpurun_out/r2_ncu_ppo_step.log 2>&1
timeout 300 python scripts/prof_own_kernels.py > gpurun_out/r2_plain_own.log 2>&1; tail -1 gpurun_out/r2_plain_own.log
for k in conv1_fwd_ws conv2_dgrad_conv1_wgrad; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r2_own_$k python scripts/prof_own_kernels.py > gpurun_out/r2_ncu_own_$k.log 2>&1; tail -1 gpurun_out/r2_ncu_own_$k.log
done
TA_CONV1_TC=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1_fwd_tc -s 2 -c 1 -f -o gpurun_out/r2_own_conv1_fwd_tc python scripts/prof_own_kernels.py > gpurun_out/r2_ncu_own_conv1_fwd_tc.log 2>&1; tail -1 gpurun_out/r2_ncu_own_conv1_fwd_tc.log
for k in conv1_bwd_tc conv2_dgrad_planes_ws; do
  TA_STEM_BWD_FUSED=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r2_own_$k python scripts/prof_own_kernels.py > gpurun_out/r2_ncu_own_$k.log 2>&1; tail -1 gpurun_out/r2_ncu_own_$k.log
done
timeout 300 python scripts/probe_dgrad.py > gpurun_out/r2_dgrad_role_profile.txt 2>&1; tail -5 gpurun_out/r2_dgrad_role_profile.txt | cut -c1-300
ls -la gpurun_out/*.ncu-rep
