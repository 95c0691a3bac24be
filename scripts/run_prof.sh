set -x
python scripts/step_time.py > gpurun_out/st_v17.log 2>&1
V=7 python scripts/step_time.py > gpurun_out/st_v7.log 2>&1
TA_DEBUG_FLAGS=1 python scripts/step_time.py > gpurun_out/st_v17_noobs.log 2>&1
TA_DEBUG_FLAGS=2 python scripts/step_time.py > gpurun_out/st_v17_noscalar.log 2>&1
TA_DEBUG_FLAGS=3 python scripts/step_time.py > gpurun_out/st_v17_nothing.log 2>&1
cat gpurun_out/st_*.log
python scripts/prof_step.py > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_obs -s 16 -c 2 -f -o gpurun_out/prof_r1d_v17_t1 python scripts/prof_step.py > gpurun_out/ncu.log 2>&1
V=7 ncu --set full --clock-control none --import-source on -k regex:step_obs -s 16 -c 2 -f -o gpurun_out/prof_r1d_v7_t1 python scripts/prof_step.py >> gpurun_out/ncu.log 2>&1
tail -3 gpurun_out/ncu.log
