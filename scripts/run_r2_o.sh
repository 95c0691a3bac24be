# full GPU suite + default bench with both data-gradient paths (tcgen05 class-major planes / cuDNN merged planes)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/r2_bench_o_tc.json 2> gpurun_out/r2_bench_o_tc.err; tail -c 2500 gpurun_out/r2_bench_o_tc.json
TA_CONV2_DGRAD_TC=0 timeout 600 python bench.py > gpurun_out/r2_bench_o_cudnn.json 2> gpurun_out/r2_bench_o_cudnn.err
python - <<'PY'
import json
for n in ("tc", "cudnn"):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_o_{n}.json").read().strip().splitlines()[-1])
        p = d["extra"]["ppo"]
        print(n, "step us/launch", d.get("ms_per_step"), "ppo value", p.get("value"), "ms/opt step", p.get("ms_per_optimizer_step"), "launches", p.get("launches_per_optimizer_step"), p.get("phases_ms"))
    except Exception as e:
        print(n, "failed", e)
PY
