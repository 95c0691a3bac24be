mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ppo.py -x -q -s 2>&1 | grep -vE "^$|Warning|warn" | tail -30
timeout 900 python bench.py --workload ppo --ppo-epochs 2 > gpurun_out/r2f_ppo.json 2> gpurun_out/r2f_ppo.err; echo "ppo rc=$?"; tail -c 1500 gpurun_out/r2f_ppo.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2f_ppo.json')); p=d['extra']['ppo']
    print('ppo',p['value'],'ms/opt step',p['ms_per_optimizer_step'],'launches',p.get('launches_per_optimizer_step'),p.get('own_launches_per_optimizer_step'),p['phases_ms'], p['losses'])
except Exception as e: print('parse fail',e)
PY
