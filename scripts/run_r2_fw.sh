timeout 900 python -m pytest tests/test_gpu_fused_kernels.py -q -x -k "conv1_fwd_ws" 2>&1 | tail -3
for d in 0 1 2 3; do echo "TA_FW_DBG=$d"; TA_FW_DBG=$d timeout 300 python scripts/probe_conv1_fwd.py 2>&1 | grep "mode 2"; done
