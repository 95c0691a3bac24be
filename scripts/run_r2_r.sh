mkdir -p gpurun_out
timeout 600 python scripts/prof_fused_timeline.py > gpurun_out/r2_fused_timeline_2streams.txt 2>&1; grep -E "replay:|activities" gpurun_out/r2_fused_timeline_2streams.txt
TWO=0 timeout 600 python scripts/prof_fused_timeline.py > gpurun_out/r2_fused_timeline_1stream.txt 2>&1; grep -E "replay:|activities" gpurun_out/r2_fused_timeline_1stream.txt
