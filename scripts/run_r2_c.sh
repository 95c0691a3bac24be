# round 2, GPU run C: full GPU tests with the new defaults, default bench line, aux kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2c_bench.err
timeout 300 python bench.py --workload aux > gpurun_out/r2c_aux.json 2> gpurun_out/r2c_aux.err; echo "aux rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2c_bench.json'))
    print('value',d['value'],'us',d['roofline']['launch_us'],'frac',d['roofline']['frac'])
    print('e2e',d['e2e']['value']); print('dma',d['extra']['e2e_dma']['value'])
    print('view7',d['extra']['view7']); print('two_streams',d['extra']['two_streams']['us_per_launch'])
    p=d['extra']['ppo']; print('ppo',p['value'],p['ms_per_optimizer_step'],p.get('launches_per_optimizer_step'),p['phases_ms'])
    a=json.load(open('gpurun_out/r2c_aux.json'))
    for k,v in a.items():
        if isinstance(v,dict): print(k, round(v['us'],2), round(v['frac'],3))
except Exception as e: print('parse fail',e)
PY
