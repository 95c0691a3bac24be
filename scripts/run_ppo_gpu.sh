python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python bench.py --workload ppo --ppo-steps 1 --ppo-warmup 1 > gpurun_out/bench_ppo.json 2> gpurun_out/bench_ppo.err; echo "ppo rc=$?"; cat gpurun_out/bench_ppo.json; tail -5 gpurun_out/bench_ppo.err
