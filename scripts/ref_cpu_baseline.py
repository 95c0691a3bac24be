"""Times the UNMODIFIED Python reference on this machine's CPU (only where /root/reference
exists; it cannot travel to the GPU box).  SURVEY.md section 8d's three figures:
  (i)   env-only:  env.step incl. gen_obs, one process, random actions
  (ii)  as-run:    Env_transact.step (incl. get_full_render) + matrix_env + data_env + stack roll + store
  (iii) PPO frames/s: the train_ppo.py loop (select_action B=1, update K=10 x 16 minibatches of 128), device=cpu
Writes profiles/r1_reference_cpu_container.json."""
import json
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import ref_shim  # noqa: E402

ref_shim.install()
tbx = types.ModuleType("tensorboardX")
tbx.SummaryWriter = type("SummaryWriter", (), {"__init__": lambda s, *a, **k: None, "add_scalar": lambda s, *a, **k: None})
sys.modules["tensorboardX"] = tbx
sys.modules["seaborn"] = types.ModuleType("seaborn")
import gym  # noqa: E402
import env_buffer  # noqa: E402
import agent.PPO as ref_ppo  # noqa: E402

ref_ppo.heatmap = lambda *a, **k: None


class NoWindow:
    def set_caption(self, *a): pass
    def show_img(self, *a): pass
    def __getattr__(self, k): return lambda *a, **kw: None


def main():
    import io, contextlib
    out = {"cpu_count": os.cpu_count(), "torch_threads": torch.get_num_threads()}
    np.random.seed(9981)
    env = gym.make("MiniGrid-twoarmy-17x17-v4", seed=9981, new_step_api=True, tile_size=17)
    acts = np.random.RandomState(7).choice(np.array([0, 1, 2, 3, 6]), size=3000)
    env.reset()
    t0 = time.perf_counter()
    for a in acts:
        _, _, te, tr, _ = env.step(int(a))
        if te or tr:
            env.reset()
    out["env_only_steps_per_s_one_process"] = len(acts) / (time.perf_counter() - t0)

    args = types.SimpleNamespace(server=True)
    window = NoWindow()
    buffer = env_buffer.Buffer_gridworld()
    buffer.grid_size = 17
    buffer.transition = np.dtype([('s', np.float32, (5, 289)), ('a', np.int64, (1,)), ('p', np.float32, (5, 2)), ('g', np.float32, (2,)),
                                  ('r', np.float32, (1,)), ('d', np.float32, (1,)), ('a_logp', np.float32, (1,))])
    buffer.buffer_capacity = 2048
    buffer.buffer = np.empty(buffer.buffer_capacity, dtype=buffer.transition)
    agent = ref_ppo.PPO()
    device = torch.device("cpu")
    sink = io.StringIO()
    frames = 0
    t_env = 0.0
    t_start = time.perf_counter()
    with contextlib.redirect_stdout(sink):
        done_update = False
        while not done_update:
            et = env_buffer.Env_transact()
            sm, ss, goal = et.reset(env, window)
            for t in range(10000):
                a_ind, a_logp = agent.select_action(sm, ss, goal, device)
                action = et.env_action(env, a_ind)
                te0 = time.perf_counter()
                _, reward, terminated, truncated, done = et.step(env, window, action, args)
                state, goal = et.data_env(env)
                ss = np.append(np.delete(ss, 0, 0), [state], 0)
                m = et.matrix_env(env)
                sm = np.append(np.delete(sm, 0, 0), [m], 0)
                buffer.store((np.array(sm, dtype='float32'), np.array([a_ind], dtype='int64'), np.array(ss, dtype='float32'),
                              np.array(goal, dtype='float32'), np.array([reward], dtype='float32'), np.array([done], dtype='int64'),
                              np.array([a_logp], dtype='float32')))
                t_env += time.perf_counter() - te0
                frames += 1
                if buffer.full:
                    tu = time.perf_counter()
                    agent.update(buffer.buffer, device, 0)
                    out["update_s"] = time.perf_counter() - tu
                    done_update = True
                    break
                if terminated or truncated:
                    break
    total = time.perf_counter() - t_start
    out["as_run_env_steps_per_s_one_process"] = frames / t_env
    out["ppo_frames_per_s"] = frames / total
    out["ppo_frames"] = frames
    out["ppo_total_s"] = total
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "profiles", "r1_reference_cpu_container.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
