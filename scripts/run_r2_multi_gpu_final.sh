# N ranks: the default bench line (step workload + extra.ppo), then PPO + predictor (configs[4]) under the same torchrun world
mkdir -p gpurun_out
N=${N:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n$N.json 2> gpurun_out/r2_final_bench_n$N.err; echo "bench N=$N rc=$?"; tail -c 600 gpurun_out/r2_final_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --workload ppo --ppo-predictor --ppo-envs 8192 --ppo-horizon 32 --ppo-epochs 2 --no-cpu-baseline > gpurun_out/r2_final_ppo_pred_n$N.json 2> gpurun_out/r2_final_ppo_pred_n$N.err; echo "ppo+predictor N=$N rc=$?"; tail -c 400 gpurun_out/r2_final_ppo_pred_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_final_bench_n$N.json").read().strip().splitlines()[-1])
    print('value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
    p=d['extra']['ppo']; print('ppo',p['value'],'ms/opt',p['ms_per_optimizer_step'],p['phases_ms'])
    d=json.loads(open("gpurun_out/r2_final_ppo_pred_n$N.json").read().strip().splitlines()[-1])
    p=d.get('extra',{}).get('ppo',d); print('ppo+predictor',d['value'],p.get('phases_ms'))
except Exception as e: print('parse fail',e)
PY
