for m in 2 1; do
echo "TA_CONV1_TC=$m"
TA_CONV1_TC=$m timeout 300 python scripts/prof_fused_timeline.py 2>&1 | grep -E "replay:|activities"
TA_CONV1_TC=$m TWO=0 timeout 300 python scripts/prof_fused_timeline.py 2>&1 | grep -E "replay:|activities"
TA_CONV1_TC=$m timeout 600 python bench.py --workload ppo --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p = d if 'phases_ms' in d else d['extra']['ppo']
print('ppo', p['value'], p['phases_ms'], p.get('graph_replayed_optimizer_steps'))"
done
