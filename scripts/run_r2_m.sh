timeout 300 python -m pytest tests/test_gpu_fused_kernels.py -q -x -k "conv2_dgrad" 2>&1 | tail -4
timeout 300 python scripts/probe_dgrad.py 2>&1 | tail -8
