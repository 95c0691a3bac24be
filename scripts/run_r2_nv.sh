timeout 600 python -m pytest tests/test_gpu_ppo.py -q -x -k "next_value or rollout_and_update or kept_across" 2>&1 | tail -8
for v in 1 0; do
TA_PPO_SHARE_NEXT_VALUE=$v TA_PPO_TIMING=1 timeout 600 python bench.py --workload ppo --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p = d if 'phases_ms' in d else d['extra']['ppo']
print('share=$v ppo', p['value'], p['phases_ms'], p.get('update_phases_ms', {}).get('prepare_ms'), p.get('losses'))"
done
