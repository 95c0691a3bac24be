mkdir -p gpurun_out
for a in 0 1; do
TA_STEP_HOST_AUTO=$a timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2953$a bench.py --gpus 8 --steps 200 --warmup 20 --no-ppo --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('auto=$a', 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['d2h_bytes_per_step'], d['e2e']['host_threads'])"
done
nproc
