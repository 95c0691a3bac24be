mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "programmatic" 2>&1 | tail -3
timeout 600 python scripts/probe_conv1_bwd_err.py 2>&1 | tail -14
