timeout 600 python -m pytest tests/test_gpu_fused_kernels.py -q -x -k "conv1_bwd_fused" 2>&1 | tail -15
timeout 300 python scripts/probe_dgrad.py 2>&1 | tail -4
