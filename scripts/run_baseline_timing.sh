set -x
python scripts/step_time.py > gpurun_out/st_v17.log 2>&1
V=7 python scripts/step_time.py > gpurun_out/st_v7.log 2>&1
TA_XPOSE=0 python scripts/step_time.py > gpurun_out/st_v17_nox.log 2>&1
python scripts/rollout_time.py > gpurun_out/rt_x1.log 2>&1
TA_XPOSE=0 python scripts/rollout_time.py > gpurun_out/rt_x0.log 2>&1
V=7 python scripts/rollout_time.py > gpurun_out/rt_v7.log 2>&1
cat gpurun_out/st_*.log gpurun_out/rt_*.log
