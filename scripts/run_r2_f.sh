mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ppo.py -q -s -k "fused_step_gradients" 2>&1 | grep -E "^(actor|critic)\." 
