mkdir -p gpurun_out
timeout 120 python scripts/prof_gae_small.py > gpurun_out/plain_gae_small.log 2>&1 && \
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"gae" -s 4 -c 1 -f -o gpurun_out/r2_gae_small python scripts/prof_gae_small.py > gpurun_out/ncu_gae_small.log 2>&1
tail -2 gpurun_out/ncu_gae_small.log
