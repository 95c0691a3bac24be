mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ppo.py tests/test_gpu_fused_kernels.py -q -x 2>&1 | tail -3
for f in 1 0; do
TA_STEM_BWD_FUSED=$f timeout 600 python scripts/prof_fused_timeline.py 2>&1 | grep -E "replay:|activities"
TA_STEM_BWD_FUSED=$f TWO=0 timeout 600 python scripts/prof_fused_timeline.py 2>&1 | grep -E "replay:|activities"
done
