#!/usr/bin/env python
"""soa/train_ppo.py's loop on N parallel B200 envs (one process per GPU).

    python scripts/train_ppo_vec.py --env MiniGrid-twoarmy-17x17-v4 --num-envs 16384 --horizon 128 --updates 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/train_ppo_vec.py --num-envs 16384 ...        # --num-envs is GLOBAL; envs are sharded by rank

Reference flow (soa/train_ppo.py:99-160): select_action -> env_action -> env.step -> matrix_env /
data_env -> frame-stack roll -> buffer.store -> update when the buffer is full.  Here every env of
the shard advances together, the buffer is [horizon, envs] on the device, and gradients are
all-reduced over NCCL (the only collective).  --her appends the hindsight relabels of every episode
that ended in the rollout (train_ppo.py:128-134 with args.her on), computed on the device."""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--env", default="MiniGrid-twoarmy-17x17-v4")       # train_ppo.py:25
    ap.add_argument("--seed", type=int, default=9981)
    ap.add_argument("--gamma", type=float, default=0.99)
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--num-envs", type=int, default=16384, help="global number of envs")
    ap.add_argument("--horizon", type=int, default=128)
    ap.add_argument("--minibatch", type=int, default=4096, help="per-rank minibatch of the update")
    ap.add_argument("--epochs", type=int, default=10, help="K_epochs (PPO.py:66)")
    ap.add_argument("--updates", type=int, default=10)
    ap.add_argument("--view", type=int, default=17)
    ap.add_argument("--fp32", action="store_true", help="no bf16 autocast")
    ap.add_argument("--her", action="store_true", help="append hindsight relabels (Buffer_gridworld.her_func) to every update")
    ap.add_argument("--predictor", default=None, nargs="?", const="",
                    help="soa/train_ppo_predictor.py: PPO + frozen frame predictor; optional checkpoint with model_encoder / model_decoder / model_predictor")
    ap.add_argument("--save", default="")
    args = ap.parse_args(argv)

    import torch
    import twoarmy_b200 as pkg
    P = importlib.import_module(pkg.__name__ + ".ppo")

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    n_local = args.num_envs // world
    torch.manual_seed(args.seed)            # same initial weights on every rank (train_ppo.py:49-56)
    if args.predictor is not None:
        agent = importlib.import_module(pkg.__name__ + ".predictor").ppo_predictor(device=dev, autocast=not args.fp32)
        if args.predictor:
            agent.load_predictor(torch.load(args.predictor, map_location=dev))
    else:
        agent = P.PPO(device=dev, autocast=not args.fp32)
    agent.gamma, agent.lr = args.gamma, args.lr
    agent.broadcast_parameters()
    torch.manual_seed(args.seed + 1000 * (rank + 1))   # action sampling differs per shard
    env = pkg.TwoarmyVecEnv(args.env, n_local, args.view, device=dev, seed=args.seed, env_id0=rank * n_local, autoreset=False)
    roll = P.VecRollout(env, agent, args.horizon)
    for u in range(args.updates):
        t0 = time.time()
        buf = roll.collect()
        torch.cuda.synchronize()
        t1 = time.time()
        data = P.with_her(buf, seed=args.seed + u, env_id0=rank * n_local, first=4 if args.predictor is not None else 0) if args.her else buf.flat()
        al, vl = agent.update(data, minibatch=args.minibatch, epochs=args.epochs)
        torch.cuda.synchronize()
        t2 = time.time()
        if rank == 0:
            frames = args.horizon * n_local * world
            print(json.dumps({"update": u, "frames": frames, "rollout_s": round(t1 - t0, 3), "update_s": round(t2 - t1, 3),
                              "frames_per_s": round(frames / (t2 - t0), 1), "mean_reward": float(buf.r.mean()),
                              "action_loss": al, "value_loss": vl}))
    if args.save and rank == 0:
        agent.save_param(args.save, args.updates)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
