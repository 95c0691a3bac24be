#!/usr/bin/env python
"""Times the UNMODIFIED Python reference on this machine's host cores.

The reference files are never copied into the repository's history: in the build container they are
read from /root/reference; `__graft_entry__.build()` places a git-ignored copy under baseline/_ref so
that it travels to the GPU box with the snapshot (bench.py's CPU legs are the only users).  The import
shim (tests/golden/ref_shim.py) stubs the absent `gym` / `matplotlib` / `turtle` packages and carries
no arithmetic.

    python baseline/ref_arm.py c1  --version 4 --view 17 --procs P --rounds K --warmup W [--round-steps S]
    python baseline/ref_arm.py ppo --version 4 --frames 256 --threads P

c1   P processes, one reference env each (`gym.make(id, agent_view_size=V)`), uniform actions over
     {0,1,2,3,6}, np.random.seed(9981 + rank), env.reset() on done; times env.step only (it includes
     gen_obs).  One "round" = every process steps S env steps; value = P*S*K / sum over timed rounds of
     the slowest process's round time.                     (gym_minigrid/envs/twoarmy_v4.py:82,
                                                            gym_minigrid/minigrid.py:1333-1441)
ppo  the loop of soa/train_ppo.py:99-160 verbatim (Env_transact.reset/step incl. get_full_render,
     matrix_env, data_env, stack roll, Buffer_gridworld.store, PPO.select_action at B = 1, PPO.update
     with K_epochs = 10 x minibatches of 128) on the CPU with torch threads = P, on a buffer of
     `--frames` records instead of 2048 (every cost in the loop is linear in the buffer size, so
     frames/s is the same; 2048 frames take ~2 minutes).   (soa/agent/PPO.py:73-161, soa/env_buffer.py)
Prints ONE JSON line.
"""
import argparse
import json
import os
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def _env_id(version):
    return f"MiniGrid-twoarmy-17x17-v{version}"


def _c1_worker(rank, version, view, rounds, round_steps, barrier, queue):
    import numpy as np
    import ref_shim
    ref_shim.install()
    import gym
    np.random.seed(9981 + rank)
    env = gym.make(_env_id(version), agent_view_size=view, seed=9981 + rank, new_step_api=True)
    acts = [int(a) for a in np.random.RandomState(7 + rank).choice(np.array([0, 1, 2, 3, 6]), size=round_steps)]
    env.reset()
    barrier.wait()
    times = []
    for _ in range(rounds):
        t0 = time.perf_counter()
        for a in acts:
            _, _, te, tr, _ = env.step(a)
            if te or tr:
                env.reset()
        times.append(time.perf_counter() - t0)
    queue.put((rank, times))


def run_c1(version, view, procs, rounds, warmup, round_steps=0):
    import multiprocessing as mp
    import numpy as np
    import ref_shim
    root = ref_shim.find_root()
    if root is None:
        return {"unavailable": "reference sources not found (/root/reference or baseline/_ref)"}
    if round_steps <= 0:  # size a round to ~0.25 s of one process
        ref_shim.install()
        import gym
        np.random.seed(1)
        env = gym.make(_env_id(version), agent_view_size=view)
        env.reset()
        t0 = time.perf_counter()
        for i in range(60):
            _, _, te, tr, _ = env.step(int((0, 1, 2, 3, 6)[i % 5]))
            if te or tr:
                env.reset()
        round_steps = max(20, int(0.25 * 60 / (time.perf_counter() - t0)))
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(procs)
    queue = ctx.Queue()
    ps = [ctx.Process(target=_c1_worker, args=(r, version, view, warmup + rounds, round_steps, barrier, queue)) for r in range(procs)]
    for p in ps:
        p.start()
    res = [queue.get() for _ in ps]
    for p in ps:
        p.join()
    per_round = [max(t[k] for _, t in res) for k in range(warmup, warmup + rounds)]
    total_s = sum(per_round)
    return {"value": procs * round_steps * rounds / total_s, "unit": "env-steps/s", "procs": procs, "round_steps": round_steps,
            "rounds": rounds, "warmup": warmup, "seconds": total_s, "ms_per_round": 1e3 * total_s / rounds,
            "per_process_steps_per_s": round_steps * rounds / (sum(sum(t[warmup:]) for _, t in res) / procs),
            "reference_root": "baseline/_ref" if root.endswith("_ref") else root}


def run_ppo(version, frames, threads):
    import contextlib
    import io
    import numpy as np
    import torch
    import ref_shim
    if ref_shim.find_root() is None:
        return {"unavailable": "reference sources not found (/root/reference or baseline/_ref)"}
    ref_shim.install()
    tbx = types.ModuleType("tensorboardX")
    tbx.SummaryWriter = type("SummaryWriter", (), {"__init__": lambda s, *a, **k: None, "add_scalar": lambda s, *a, **k: None})
    sys.modules["tensorboardX"] = tbx
    sys.modules["seaborn"] = types.ModuleType("seaborn")
    import gym
    import env_buffer
    import agent.PPO as ref_ppo
    ref_ppo.heatmap = lambda *a, **k: None          # plotting side effect with hard-coded paths (PPO.py:161)
    torch.set_num_threads(threads)

    class NoWindow:                                   # gym_minigrid/window.py is a matplotlib UI
        def __getattr__(self, k):
            return lambda *a, **kw: None

    np.random.seed(9981)
    torch.manual_seed(9981)
    env = gym.make(_env_id(version), seed=9981, new_step_api=True, tile_size=17)
    args = types.SimpleNamespace(server=True)
    window = NoWindow()
    buffer = env_buffer.Buffer_gridworld()
    buffer.grid_size = 17
    buffer.transition = np.dtype([('s', np.float32, (5, 289)), ('a', np.int64, (1,)), ('p', np.float32, (5, 2)), ('g', np.float32, (2,)),
                                  ('r', np.float32, (1,)), ('d', np.float32, (1,)), ('a_logp', np.float32, (1,))])   # train_ppo.py:93-97
    buffer.buffer_capacity = frames
    buffer.buffer = np.empty(buffer.buffer_capacity, dtype=buffer.transition)
    agent = ref_ppo.PPO()
    device = torch.device("cpu")
    n, t_env, t_act, t_upd = 0, 0.0, 0.0, 0.0
    t_start = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        done_update = False
        while not done_update:
            et = env_buffer.Env_transact()
            sm, ss, goal = et.reset(env, window)
            for _ in range(10000):
                ta = time.perf_counter()
                a_ind, a_logp = agent.select_action(sm, ss, goal, device)
                te0 = time.perf_counter()
                t_act += te0 - ta
                action = et.env_action(env, a_ind)
                _, reward, terminated, truncated, done = et.step(env, window, action, args)
                state, goal = et.data_env(env)
                ss = np.append(np.delete(ss, 0, 0), [state], 0)
                m = et.matrix_env(env)
                sm = np.append(np.delete(sm, 0, 0), [m], 0)
                buffer.store((np.array(sm, dtype='float32'), np.array([a_ind], dtype='int64'), np.array(ss, dtype='float32'),
                              np.array(goal, dtype='float32'), np.array([reward], dtype='float32'), np.array([done], dtype='int64'),
                              np.array([a_logp], dtype='float32')))
                t_env += time.perf_counter() - te0
                n += 1
                if buffer.full:
                    tu = time.perf_counter()
                    agent.update(buffer.buffer, device, 0)
                    t_upd = time.perf_counter() - tu
                    done_update = True
                    break
                if terminated or truncated:
                    break
    total = time.perf_counter() - t_start
    return {"value": n / total, "unit": "frames/s", "frames": n, "seconds": total, "threads": threads,
            "rollout_env_s": t_env, "select_action_s": t_act, "update_s": t_upd,
            "as_run_env_steps_per_s_one_process": n / t_env, "K_epochs": int(agent.K_epochs), "minibatch": int(agent.batch_size)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["c1", "ppo"])
    ap.add_argument("--version", type=int, default=4)
    ap.add_argument("--view", type=int, default=17)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--rounds", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--round-steps", type=int, default=0)
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    if a.mode == "c1":
        out = run_c1(a.version, a.view, a.procs, a.rounds, a.warmup, a.round_steps)
    else:
        out = run_ppo(a.version, a.frames, a.threads)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
