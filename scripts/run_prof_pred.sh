mkdir -p gpurun_out
timeout 120 python scripts/prof_pred_kernels.py > gpurun_out/plain_pred.log 2>&1 && \
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"pred_(en|de)coder" -s 2 -c 2 -f -o gpurun_out/r2_pred_kernels python scripts/prof_pred_kernels.py > gpurun_out/ncu_pred.log 2>&1
tail -2 gpurun_out/ncu_pred.log
