timeout 300 python -m pytest tests/test_gpu_fused_kernels.py -q -x -k "conv1_fwd_ws" 2>&1 | tail -3
timeout 120 python scripts/probe_conv1_fwd.py 2>&1 | grep "mode 2"
