# round 2, GPU run D: full GPU tests, GAE small-rollout variants, bench with the PDL default
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25 > gpurun_out/r2d_pytest.txt; tail -8 gpurun_out/r2d_pytest.txt
timeout 300 python -m pytest tests/test_gpu_ppo.py -q -s -k "fixture or graph_replay" 2>&1 | grep -E "cosine|bit-identical|passed|failed" 
for var in 0 1 2 3; do echo "== TA_GAE_SMALL=$var"; TA_GAE_SMALL=$var timeout 300 python bench.py --workload aux 2>/dev/null | python -c "
import json,sys; a=json.loads(sys.stdin.read())
print({k:(round(v['us'],2), round(v['frac'],3)) for k,v in a.items() if k.startswith('gae')})"; done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-ppo > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2d_bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2d_bench.json'))
    print('value',d['value'],'us',d['roofline']['launch_us'],'frac',d['roofline']['frac'])
    print('e2e',d['e2e']['value']); print('view7',d['extra']['view7']); print('two_streams',d['extra']['two_streams']['us_per_launch'])
    print('launch_size',d['extra']['launch_size']); print('rollout', d['extra']['rollout'])
except Exception as e: print('parse fail',e)
PY
timeout 300 python bench.py --gpus 1 --steps 2000 --warmup 200 --no-ppo --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('K=2000: us',d['roofline']['launch_us'],'frac',d['roofline']['frac'],'view7',d['extra']['view7']['us_per_launch'])"
