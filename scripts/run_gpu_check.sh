# GPU parity tests, then T=1 and rollout timings of the step kernel (device-timed)
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
python scripts/step_time.py > gpurun_out/st_v17.log 2>&1
V=7 python scripts/step_time.py > gpurun_out/st_v7.log 2>&1
python scripts/rollout_time.py > gpurun_out/rt_v17.log 2>&1
V=7 python scripts/rollout_time.py > gpurun_out/rt_v7.log 2>&1
cat gpurun_out/st_v17.log gpurun_out/st_v7.log gpurun_out/rt_v17.log gpurun_out/rt_v7.log
