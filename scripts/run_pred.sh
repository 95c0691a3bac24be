mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ppo.py -q -m gpu -x -k "lstm or predict or stem or fused8 or fused_step" -s 2>&1 | grep -E "passed|failed|Error|error|fused vs fp32|losses|assert" | cut -c1-220 | tail -60
timeout 900 python bench.py --workload ppo --ppo-predictor --ppo-envs 8192 --ppo-horizon 32 --ppo-epochs 2 --no-cpu-baseline > gpurun_out/bench_ppo_pred_new.json 2> gpurun_out/bench_ppo_pred_new.err; echo rc=$?
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_ppo_pred_new.json").read().strip().splitlines()[-1])
p = d.get("extra", {}).get("ppo", d)
print(d["value"], p.get("phases_ms"), p.get("graph_replayed_optimizer_steps"), p.get("fused_step"), p.get("launches_per_optimizer_step"))
PY
tail -3 gpurun_out/bench_ppo_pred_new.err | cut -c1-300
