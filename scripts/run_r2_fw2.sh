mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1_fwd_ws -s 4 -c 1 -f -o gpurun_out/r2_conv1_fwd_ws python scripts/probe_conv1_fwd.py > gpurun_out/r2_ncu_fw.log 2>&1; tail -2 gpurun_out/r2_ncu_fw.log
