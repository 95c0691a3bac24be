mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_kernels.py tests/test_gpu_ppo.py -q 2>&1 | tail -12
timeout 900 python bench.py --workload ppo --ppo-epochs 3 > gpurun_out/r2h_ppo.json 2> gpurun_out/r2h_ppo.err; echo "ppo rc=$?"; tail -c 300 gpurun_out/r2h_ppo.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2h_ppo.json')); p=d['extra']['ppo']
    print('ppo',p['value'],'ms/opt step',p['ms_per_optimizer_step'],'replayed',p.get('graph_replayed_optimizer_steps'),'launches',p.get('launches_per_optimizer_step'),p.get('own_launches_per_optimizer_step'),p['phases_ms'], p['losses'], p.get('kernel_time_per_optimizer_step_us'))
except Exception as e: print('parse fail',e)
PY
