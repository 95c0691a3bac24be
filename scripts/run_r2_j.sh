mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "vec_rollout" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_ppo.py tests/test_gpu_fused_kernels.py -q 2>&1 | tail -5
timeout 900 python bench.py --workload ppo --ppo-epochs 3 > gpurun_out/r2j_ppo.json 2> gpurun_out/r2j_ppo.err; echo "ppo rc=$?"; tail -c 600 gpurun_out/r2j_ppo.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2j_ppo.json')); p=d['extra']['ppo']
    print('ppo',p['value'],'ms/opt step',p['ms_per_optimizer_step'],'replayed',p.get('graph_replayed_optimizer_steps'),'launches',p.get('launches_per_optimizer_step'),p['phases_ms'], p['losses'])
except Exception as e: print('parse fail',e)
PY
