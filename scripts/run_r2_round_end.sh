# full GPU suite, smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_round_end.json 2> gpurun_out/r2_bench_round_end.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_round_end.json").read().strip().splitlines()[-1])
p = d["extra"]["ppo"]
print("step us", d["ms_per_step"] * 1e3, "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "view7", d["extra"]["view7"]["frac"], "cpu", d["cpu_baseline"]["value"])
print("ppo", p["value"], p["phases_ms"], p.get("graph_replayed_optimizer_steps"), p.get("launches_per_optimizer_step"), p.get("own_launches_per_optimizer_step"), p["cpu_baseline"]["value"])
PY
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_round_end_ref.json 2> gpurun_out/r2_bench_round_end_ref.err; head -c 700 gpurun_out/r2_bench_round_end_ref.json
