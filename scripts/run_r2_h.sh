timeout 900 python -m pytest tests/test_gpu_fused_kernels.py -q -k adam 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_ppo.py -q -s -k "fused_step_gradients" 2>&1 | grep -E "^(actor|critic)\.|passed|failed|Error" | head -80
