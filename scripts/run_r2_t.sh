mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ppo.py tests/test_gpu_fused_kernels.py -q -x 2>&1 | tail -3
TWO=0 timeout 600 python scripts/prof_fused_timeline.py > gpurun_out/r2_fused_timeline_1stream.txt 2>&1; grep -E "replay:|activities|relu_bwd" gpurun_out/r2_fused_timeline_1stream.txt | cut -c1-100
