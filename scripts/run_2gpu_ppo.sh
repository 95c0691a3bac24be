# 2-GPU PPO loop: eager optimiser steps with NCCL all-reduce vs the step captured in a CUDA graph including the all-reduce
for G in 0 1; do
TA_PPO_GRAPH_NCCL=$G timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29518+G)) bench.py --gpus 2 --workload ppo --ppo-envs 16384 --ppo-horizon 32 --ppo-epochs 4 --ppo-steps 1 --ppo-warmup 1 > gpurun_out/bench_ppo_n2_g$G.json 2> gpurun_out/bench_ppo_n2_g$G.err; echo "ppo n2 graph_nccl=$G rc=$?"; grep -o '"value": [0-9.]*' gpurun_out/bench_ppo_n2_g$G.json | head -1; grep -o '"phases[^}]*}' gpurun_out/bench_ppo_n2_g$G.json; grep -o '"losses[^}]*}' gpurun_out/bench_ppo_n2_g$G.json; tail -2 gpurun_out/bench_ppo_n2_g$G.err | cut -c1-200
done
