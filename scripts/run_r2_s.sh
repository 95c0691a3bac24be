mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ppo.py tests/test_gpu_fused_kernels.py -q -x 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "step_host or packed" 2>&1 | tail -3
TA_PPO_TIMING=1 timeout 600 python bench.py --workload ppo --no-cpu-baseline > gpurun_out/r2_s_ppo.json 2> gpurun_out/r2_s_ppo.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_s_ppo.json").read().strip().splitlines()[-1])
p = d if "phases_ms" in d else d["extra"]["ppo"]
print(p.get("value"), p.get("phases_ms"), p.get("graph_replayed_optimizer_steps"), p.get("update_phases_ms"))
PY
TA_HOST_THREADS=3 timeout 300 python bench.py --no-ppo --no-extra --no-cpu-baseline --steps 200 --warmup 20 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('3 threads e2e', d['e2e'])"
