python scripts/write_bw.py > gpurun_out/write_bw.log 2>&1; cat gpurun_out/write_bw.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
python bench.py --steps 64 --warmup 8 --no-extra --no-cpu-baseline --e2e-steps 2 > gpurun_out/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 64 --warmup 8 --no-extra --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_launches.log 2>&1
python scripts/prof_step.py > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_obs -s 16 -c 2 -f -o gpurun_out/prof_r1e_v17_t1 python scripts/prof_step.py > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
