mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ppo.py tests/test_gpu_fused_kernels.py -q -x 2>&1 | tail -3
TWO=0 timeout 600 python scripts/prof_fused_timeline.py > gpurun_out/r2_fused_timeline_1stream.txt 2>&1; grep -E "replay:|activities|conv1_|wgrad_kernel" gpurun_out/r2_fused_timeline_1stream.txt | cut -c1-110
timeout 600 python scripts/prof_fused_timeline.py > gpurun_out/r2_fused_timeline_2streams.txt 2>&1; grep -E "replay:|activities" gpurun_out/r2_fused_timeline_2streams.txt
