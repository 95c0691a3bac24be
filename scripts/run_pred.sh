mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ppo.py -q -m gpu -x -k "lstm or predict" 2>&1 | tail -5
timeout 300 python scripts/prof_pred_rollout.py > gpurun_out/pred_rollout_prof.txt 2>&1; cut -c1-200 gpurun_out/pred_rollout_prof.txt | grep -v "^-" | head -12
