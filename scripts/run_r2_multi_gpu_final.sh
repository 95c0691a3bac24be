# N ranks: the default bench line (step workload + extra.ppo + extra.ppo_predictor) under torchrun, then the reference arm
mkdir -p gpurun_out
N=${N:-2}
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n$N.json 2> gpurun_out/r2_final_bench_n$N.err ) 2>&1 | tail -3; echo "bench N=$N rc=$?"; tail -c 400 gpurun_out/r2_final_bench_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_final_bench_n$N.json").read().strip().splitlines()[-1])
    print('value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
    p=d['extra']['ppo']; print('ppo',p['value'],'ms/opt',p['ms_per_optimizer_step'],p['phases_ms'],p.get('allreduce_us_per_optimizer_step'))
    q=d['extra']['ppo_predictor']; print('ppo+predictor',q['value'],q['phases_ms'])
except Exception as e: print('parse fail',e)
PY
