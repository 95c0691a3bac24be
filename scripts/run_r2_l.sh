mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ppo.py tests/test_gpu_fused_kernels.py -q -x 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "stack or vec_rollout or matrix_env" 2>&1 | tail -3
timeout 300 python scripts/probe_conv1_bwd_err.py 2>&1 | tail -4
TWO=0 timeout 600 python scripts/prof_fused_timeline.py > gpurun_out/r2_fused_timeline_1stream.txt 2>&1; grep -E "replay:|activities|conv1_" gpurun_out/r2_fused_timeline_1stream.txt | head -8
timeout 600 python scripts/prof_fused_timeline.py 2>&1 | grep -E "replay:|activities"
timeout 300 python bench.py --workload aux 2>/dev/null | python -c "
import json,sys; a=json.loads(sys.stdin.read())
print({k:(round(v['us'],2), round(v['frac'],3)) for k,v in a.items() if isinstance(v,dict)})"
