mkdir -p gpurun_out
timeout 200 python scripts/prof_final_aux.py > gpurun_out/plain_final_aux.log 2>&1 && \
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"stack_push|frame_codes|gae_vec4|adv_|pred_|lstm_gates" --csv --log-file gpurun_out/r2_final_aux_kernels_ncu.csv python scripts/prof_final_aux.py > gpurun_out/ncu_final_aux.log 2>&1
tail -1 gpurun_out/ncu_final_aux.log; wc -l gpurun_out/r2_final_aux_kernels_ncu.csv
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"stack_push_tma" -s 2 -c 1 -f -o gpurun_out/r2_stack_push_tma python scripts/prof_final_aux.py > gpurun_out/ncu_push_tma.log 2>&1; tail -1 gpurun_out/ncu_push_tma.log
