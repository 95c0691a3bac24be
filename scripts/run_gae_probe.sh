mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gae.py -q -m gpu -x 2>&1 | tail -2
for v in 1 2 3; do TA_GAE_SMALL=$v timeout 120 python scripts/probe_aux_floor.py 2>/dev/null | tail -1; done | tee gpurun_out/gae_probe.txt
