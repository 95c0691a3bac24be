mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv2_dgrad_planes_ws -s 1 -c 1 -f -o gpurun_out/r2_dgrad_ws python scripts/prof_own_kernels.py > gpurun_out/r2_ncu_dgrad.log 2>&1; tail -2 gpurun_out/r2_ncu_dgrad.log
