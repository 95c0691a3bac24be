# full GPU test pass + the PPO loop bench line (BASELINE configs[3])
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 600 python bench.py --workload ppo --ppo-steps 2 --ppo-warmup 1 2>gpurun_out/bench_ppo_err.log | tail -1 > gpurun_out/bench_ppo_tc.json; cut -c1-600 gpurun_out/bench_ppo_tc.json
