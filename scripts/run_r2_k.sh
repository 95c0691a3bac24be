mkdir -p gpurun_out
timeout 300 python scripts/prof_own_kernels.py > gpurun_out/r2_plain_own.log 2>&1; tail -1 gpurun_out/r2_plain_own.log
for k in stack_push_codes_tile frame_codes_tile conv1_fwd_tc conv1_bwd_tc; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r2_own_$k python scripts/prof_own_kernels.py > gpurun_out/r2_ncu_own_$k.log 2>&1; tail -1 gpurun_out/r2_ncu_own_$k.log
done
ls -la gpurun_out/*.ncu-rep
