# round 2, GPU run B: packed host path tests, PDL variants of the step kernel (timing + parity), bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "step_host or step_packed or trajectories or rollout_equals or sharding or persistent" 2>&1 | tail -4
for pdl in 0 1 2 3; do
  for v in 17 7; do
    echo "== TA_PDL=$pdl V=$v"; TA_PDL=$pdl V=$v STEPS=1600 timeout 120 python scripts/step_time.py 2>&1 | tail -1
  done
done
for pdl in 2 3; do
  echo "== parity with TA_PDL=$pdl"; TA_PDL=$pdl timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "trajectories or rollout_equals or sharding or persistent or long_run or step_host" 2>&1 | tail -2
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-ppo > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2b_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2b_bench.json'))
    print('value',d['value'],'us',d['roofline']['launch_us'],'frac',d['roofline']['frac'])
    print('e2e',d['e2e']); print('dma',d['extra']['e2e_dma'])
    print('view7',d['extra']['view7'])
    print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['port']['value'])
except Exception as e: print('parse fail',e)
PY
for th in 4 8 16; do echo "== TA_HOST_THREADS=$th"; TA_HOST_THREADS=$th timeout 300 python bench.py --steps 20 --warmup 5 --no-ppo --no-extra --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['e2e']['value'], d['extra']['e2e_dma']['value'])"; done
