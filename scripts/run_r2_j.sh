mkdir -p gpurun_out
timeout 900 python bench.py --workload ppo --ppo-epochs 3 > gpurun_out/r2j_ppo.json 2> gpurun_out/r2j_ppo.err; echo "ppo rc=$?"; tail -c 300 gpurun_out/r2j_ppo.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2j_ppo.json')); p=d['extra']['ppo']
    print('ppo',p['value'],'ms/opt step',p['ms_per_optimizer_step'],'replayed',p.get('graph_replayed_optimizer_steps'),'launches',p.get('launches_per_optimizer_step'),p['phases_ms'], p['losses'])
except Exception as e: print('parse fail',e)
PY
TA_ROLLOUT_GRAPH=0 timeout 900 python bench.py --workload ppo --ppo-epochs 3 2>/dev/null | python -c "
import json,sys; p=json.loads(sys.stdin.read())['extra']['ppo']; print('eager rollout:',p['value'],p['phases_ms'])"
