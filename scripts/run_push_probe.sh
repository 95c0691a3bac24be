# stack_push: parity tests, then time per launch for the register kernel and the pipelined kernel's stage / occupancy settings
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "stack or push or roll or rollout" 2>&1 | tail -3
for cfg in "TA_PUSH_TMA=0 TA_PUSH_MINB=5" "TA_PUSH_STAGES=3" "TA_PUSH_STAGES=2" "TA_PUSH_STAGES=4" "TA_PUSH_STAGES=3 TA_PUSH_CTAS=2" \
           "TA_PUSH_STAGES=4 TA_PUSH_CTAS=2"; do
  env $cfg timeout 120 python scripts/probe_push.py 2>&1 | tail -1 | cut -c1-200
done | tee gpurun_out/push_probe.txt
