mkdir -p gpurun_out
timeout 600 python scripts/prof_fused_timeline.py > gpurun_out/r2_fused_timeline.txt 2>&1; tail -120 gpurun_out/r2_fused_timeline.txt
TWO=0 timeout 600 python scripts/prof_fused_timeline.py 2>&1 | grep -E "replay:|activities"
