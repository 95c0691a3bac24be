mkdir -p gpurun_out
for v in 1 0; do
TA_PPO_TIMING=1 TA_CONV2_DGRAD_TC=$v timeout 600 python bench.py --workload ppo --no-cpu-baseline > gpurun_out/r2_p_$v.json 2> gpurun_out/r2_p_$v.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_p_$v.json").read().strip().splitlines()[-1])
p = d if "phases_ms" in d else d["extra"]["ppo"]
print("tc=$v", p.get("value"), p.get("phases_ms"), p.get("graph_replayed_optimizer_steps"), p.get("update_phases_ms"))
PY
done
