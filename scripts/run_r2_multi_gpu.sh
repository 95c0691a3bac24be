mkdir -p gpurun_out
N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench N=$N rc=$?"; tail -c 1500 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
    print('value',d['value'],'us',d['roofline']['launch_us'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'], 'dma', d['extra']['e2e_dma']['value'])
    p=d['extra']['ppo']; print('ppo',p['value'],'ms/opt',p['ms_per_optimizer_step'],p.get('graph_replayed_optimizer_steps'),p.get('launches_per_optimizer_step'),p['phases_ms'],p.get('allreduce_us_per_optimizer_step'), p['config']['optimizer_steps_per_iteration'])
except Exception as e: print('parse fail',e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 20 --warmup 5 | head -c 600
