mkdir -p gpurun_out
timeout 900 python bench.py --workload ppo --ppo-predictor --ppo-envs 8192 --ppo-horizon 32 --ppo-epochs 2 --no-cpu-baseline > gpurun_out/bench_ppo_pred_new.json 2> gpurun_out/bench_ppo_pred_new.err; echo rc=$?
timeout 900 python bench.py --workload ppo --ppo-predictor --ppo-envs 16384 --ppo-horizon 128 --ppo-epochs 10 --no-cpu-baseline > gpurun_out/bench_ppo_pred_cfg3shape.json 2> gpurun_out/bench_ppo_pred_cfg3shape.err; echo rc=$?
python - <<'PY'
import json
for f in ("gpurun_out/bench_ppo_pred_new.json", "gpurun_out/bench_ppo_pred_cfg3shape.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    p = d.get("extra", {}).get("ppo", d)
    print(f, d["value"], p.get("phases_ms"), p.get("graph_replayed_optimizer_steps"))
PY
