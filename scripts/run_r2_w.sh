mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv2_dgrad_conv1_wgrad -s 3 -c 1 -f -o gpurun_out/r2_stem_bwd python scripts/probe_dgrad.py > gpurun_out/r2_ncu_stem_bwd.log 2>&1; tail -2 gpurun_out/r2_ncu_stem_bwd.log
