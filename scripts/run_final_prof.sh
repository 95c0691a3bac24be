set -x
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; echo "ref rc=$?"
timeout 300 python bench.py --workload aux > gpurun_out/aux.json 2>> gpurun_out/bench.err
timeout 200 python bench.py --steps 64 --warmup 8 --no-extra --no-cpu-baseline --e2e-steps 2 > gpurun_out/plain_launches.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 64 --warmup 8 --no-extra --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_launches.log 2>&1
timeout 200 python scripts/prof_step.py > gpurun_out/plain.log 2>&1 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:step_obs -s 16 -c 2 -f -o gpurun_out/prof_r1f_v17_t1 python scripts/prof_step.py > gpurun_out/ncu.log 2>&1
K=40 timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:step_obs -s 16 -c 16 --csv --log-file gpurun_out/traffic_steady.csv python scripts/prof_step.py > gpurun_out/ncu_traffic.log 2>&1
tail -2 gpurun_out/ncu.log; cat gpurun_out/bench.json
