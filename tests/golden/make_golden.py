"""Generates the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF
(/root/reference, imported through ref_shim.py).  Run it only where the reference exists:

    python tests/golden/make_golden.py

Outputs (committed):
    traj_v4.npz, traj_v6.npz   N envs x T steps: actions, forced resets, recorded
                               np.random.choice draws per call site, and after every step the
                               reference's reward, terminated, truncated, agent_pos, grid,
                               every Twoarmy flag, all ball positions, obs (V=17 and V=7),
                               matrix_env / data_env features
    kat_v6_goal.npz            SURVEY.md section-4 scripted known-answer trajectory
    obs_general.npz            gen_obs for agent_dir 0..3, several view sizes, both
                               see_through_walls modes (Grid.process_vis path)
    td_adv.npz                 PPO.py:112-115 arithmetic on fp32 tensors

Nothing here is imported by the product or by the GPU-side tests; they only read the .npz.
"""
import inspect
import os
import sys
from collections import deque

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import gym  # noqa: E402  (the stub)
import env_buffer  # noqa: E402  (reference soa/env_buffer.py)

CODE = {None: 0, "wall": 1, "ball": 2, "goal": 3}
REWARD_LUT = [-0.01, -0.1, -0.9, 0.2, 0.9]
# np.random.choice call sites -> slot
SLOT_BY_LINE = {
    4: {117: 0, 149: 0, 184: 1, 190: 2, 215: 3, 221: 4, 303: 5, 310: 6},
    6: {306: 5, 313: 6},
}


class DrawRecorder:
    """Wraps np.random.choice; records (slot, value) for calls made from the env files."""

    def __init__(self, version):
        self.version = version
        self.real = np.random.choice
        self.cur = None

    def __enter__(self):
        def wrapped(a, size=None, *args, **kw):
            out = self.real(a, size, *args, **kw)
            fr = inspect.currentframe().f_back
            fn = os.path.basename(fr.f_code.co_filename)
            if fn.startswith("twoarmy_v") and self.cur is not None:
                slot = SLOT_BY_LINE[self.version][fr.f_lineno]
                assert self.cur[slot] == 0xFF, "slot drawn twice in one step"
                self.cur[slot] = int(np.asarray(out).item())
            return out

        np.random.choice = wrapped
        return self

    def __exit__(self, *a):
        np.random.choice = self.real

    def begin_step(self):
        self.cur = np.full(8, 0xFF, dtype=np.uint8)

    def end_step(self):
        out, self.cur = self.cur, None
        return out


def grid_codes(env):
    return np.array([CODE[None if c is None else c.type] for c in env.grid.grid], dtype=np.uint8)


def ball_positions(env):
    out = np.full((10, 2), -1, dtype=np.int8)
    objs = list(env.obstacles) + list(env.obstacles1) + list(env.obstacles2)
    for k, o in enumerate(objs):
        if o.cur_pos is not None:
            out[k] = o.cur_pos
    return out


def flags(env):
    return np.array([env.step_count, env.step_move, env.risk_count, env.pone, env.patrol, env.up1,
                     env.right2, env.Update_horizontal, env.Update_longitudinal, env.first_to_room2],
                    dtype=np.int32)


def bfs_action(env, target, rng):
    """First move of a shortest path over empty/goal cells to `target` (None if unreachable)."""
    W = 17
    start = tuple(env.agent_pos)
    if start == target:
        return 6
    prev = {start: None}
    dq = deque([start])
    moves = [(0, (-1, 0)), (1, (1, 0)), (2, (0, -1)), (3, (0, 1))]
    while dq:
        p = dq.popleft()
        if p == target:
            break
        order = list(moves)
        rng.shuffle(order)
        for a, (dx, dy) in order:
            q = (p[0] + dx, p[1] + dy)
            if not (0 <= q[0] < W and 0 <= q[1] < W) or q in prev:
                continue
            c = env.grid.get(*q)
            if c is None or c.type == "goal":
                prev[q] = (p, a)
                dq.append(q)
    if target not in prev:
        return None
    p = target
    a = 6
    while prev[p] is not None:
        p, a = prev[p]
    return a


def pick_action(env, kind, rng, t):
    """Behaviour policies chosen to cover every branch of Twoarmy.step."""
    r = rng.random()
    if kind == 0:  # uniform over the five actions the adapter can emit
        return int(rng.choice([0, 1, 2, 3, 6]))
    if kind == 1 or kind == 2:  # noisy shortest path to the goal
        eps = 0.3 if kind == 1 else 0.1
        if r < eps:
            return int(rng.choice([0, 1, 2, 3, 6]))
        a = bfs_action(env, tuple(env.goal_pos), rng)
        return int(rng.choice([0, 1, 2, 3, 6])) if a is None else a
    if kind == 3:  # loiter under the oscillating balls (risk_count path), some clamped actions
        if r < 0.1:
            return int(rng.choice([7, 9, 100]))  # >= 7 -> clamped to 0 (twoarmy_v4.py:84-85)
        a = bfs_action(env, (8, 9), rng)
        return 6 if a is None else a
    if kind == 4:  # walk into the mid balls / patrols
        if r < 0.2:
            return int(rng.choice([0, 1, 2, 3, 6]))
        tgt = (8, 8) if tuple(env.agent_pos)[1] > 8 else (9, 5)
        a = bfs_action(env, tgt, rng)
        if a is None:  # target occupied: step towards it anyway
            ax, ay = env.agent_pos
            a = 2 if ay > tgt[1] else (1 if ax < tgt[0] else (0 if ax > tgt[0] else 3))
        return a
    raise ValueError(kind)


def run_traj(version, N, T, view, seed0, with_forced_resets=True):
    env_id = f"MiniGrid-twoarmy-17x17-v{version}"
    out = dict(
        actions=np.zeros((T, N), np.int32), forced_reset=np.zeros((T, N), np.uint8),
        draws=np.full((T, N, 8), 0xFF, np.uint8), reward_idx=np.zeros((T, N), np.uint8),
        reward=np.zeros((T, N), np.float64), term=np.zeros((T, N), np.uint8),
        trunc=np.zeros((T, N), np.uint8), agent=np.zeros((T, N, 2), np.int8),
        grid=np.zeros((T, N, 289), np.uint8), flags=np.zeros((T, N, 10), np.int32),
        balls=np.zeros((T, N, 10, 2), np.int8), obs=np.zeros((T, N, view, view, 3), np.uint8),
        reset_obs=np.zeros((N, view, view, 3), np.uint8),
        post_reset_obs=np.zeros((T, N, view, view, 3), np.uint8),
        matrix=np.zeros((T, N, 289), np.float32), place=np.zeros((T, N, 2), np.float32),
        goal=np.zeros((N, 2), np.float32), kind=np.zeros(N, np.int32),
    )
    for n in range(N):
        np.random.seed(seed0 + n)  # the env's RNG is the global legacy stream (train_ppo.py:51)
        rng = np.random.default_rng(1000 + n)  # behaviour policy: separate generator
        kind = n % 5
        out["kind"][n] = kind
        env = gym.make(env_id, agent_view_size=view)
        et = env_buffer.Env_transact()
        out["reset_obs"][n] = env.reset()["image"]
        out["goal"][n] = et.data_env(env)[1]
        with DrawRecorder(version) as rec:
            for t in range(T):
                a = pick_action(env, kind, rng, t)
                out["actions"][t, n] = a
                rec.begin_step()
                obs, r, te, tr, _ = env.step(a)
                out["draws"][t, n] = rec.end_step()
                ri = REWARD_LUT.index(r)  # the reference only returns these literals
                out["reward_idx"][t, n] = ri
                out["reward"][t, n] = r
                out["term"][t, n], out["trunc"][t, n] = te, tr
                out["agent"][t, n] = env.agent_pos
                out["grid"][t, n] = grid_codes(env)
                out["flags"][t, n] = flags(env)
                out["balls"][t, n] = ball_positions(env)
                out["obs"][t, n] = obs["image"]
                out["matrix"][t, n] = et.matrix_env(env).astype(np.float32)
                out["place"][t, n] = et.data_env(env)[0]
                if te or tr:
                    out["post_reset_obs"][t, n] = env.reset()["image"]
                elif with_forced_resets and rng.random() < 0.01 and not env.patrol:
                    # mid-episode reset(): flags carry over (minigrid.py:947-980). v4 with
                    # patrol=True would crash on the next step, so it is not exercised here.
                    out["forced_reset"][t, n] = 1
                    out["post_reset_obs"][t, n] = env.reset()["image"]
    return out


def make_traj(version, N=40, T=160):
    a = run_traj(version, N, T, 17, 9981)
    b = run_traj(version, N, T, 7, 9981)
    for k in a:
        if k not in ("obs", "reset_obs", "post_reset_obs"):
            assert np.array_equal(a[k], b[k]), k  # the view size must not change the dynamics
    a["obs7"], a["reset_obs7"], a["post_reset_obs7"] = b["obs"], b["reset_obs"], b["post_reset_obs"]
    a["obs17"], a["reset_obs17"], a["post_reset_obs17"] = a.pop("obs"), a.pop("reset_obs"), a.pop("post_reset_obs")
    np.savez_compressed(os.path.join(HERE, f"traj_v{version}.npz"), **a)
    d = a["draws"]
    print(f"v{version}: episodes={int((a['term'] | a['trunc']).sum())} goals={int(a['term'].sum())} "
          f"hits={int((a['reward_idx'] == 2).sum())} risk={int((a['reward_idx'] == 1).sum())} "
          f"room2={int((a['reward_idx'] == 3).sum())} patrol_steps={int(a['flags'][..., 4].sum())} "
          f"forced={int(a['forced_reset'].sum())} draws/slot={[(int((d[..., s] != 255).sum())) for s in range(7)]}")


def make_kat():
    env = gym.make("MiniGrid-twoarmy-17x17-v6")
    np.random.seed(1)
    env.reset()
    acts = [2] * 6 + [1] * 7 + [2] * 7 + [1] * 4
    rew, pos, te_, tr_ = [], [], [], []
    for a in acts:
        _, r, te, tr, _ = env.step(a)
        rew.append(r); pos.append(env.agent_pos); te_.append(te); tr_.append(tr)
    assert te_[-1] and tuple(env.agent_pos) == (14, 2)
    np.savez_compressed(os.path.join(HERE, "kat_v6_goal.npz"), actions=np.array(acts, np.int32),
                        reward=np.array(rew, np.float64), agent=np.array(pos, np.int8),
                        term=np.array(te_, np.uint8), trunc=np.array(tr_, np.uint8))
    print("kat rewards", rew)


def make_obs_general():
    """gen_obs with agent_dir / see_through_walls overridden on a live reference env
    (Twoarmy fixes them, so the general path is exercised by poking the attributes)."""
    rng = np.random.default_rng(5)
    np.random.seed(3)
    rows = []
    for version in (4, 6):
        env = gym.make(f"MiniGrid-twoarmy-17x17-v{version}")
        env.reset()
        for t in range(400):
            a = pick_action(env, 1, rng, t)
            _, _, te, tr, _ = env.step(a)
            if t % 4 == 0:
                for V in (3, 5, 7, 9, 17):
                    d = int(rng.integers(4)); stw = bool(rng.integers(2))
                    env.agent_view_size, env.agent_dir, env.see_through_walls = V, d, stw
                    img = env.gen_obs()["image"]
                    pad = np.zeros((17, 17, 3), np.uint8); pad[:V, :V] = img
                    rows.append((grid_codes(env), env.agent_pos[0], env.agent_pos[1], d, V, int(stw), pad))
                env.agent_view_size, env.agent_dir, env.see_through_walls = 17, 3, True
            if te or tr:
                env.reset()
    np.savez_compressed(
        os.path.join(HERE, "obs_general.npz"),
        grid=np.stack([r[0] for r in rows]), ax=np.array([r[1] for r in rows], np.int32),
        ay=np.array([r[2] for r in rows], np.int32), dir=np.array([r[3] for r in rows], np.int32),
        view=np.array([r[4] for r in rows], np.int32), stw=np.array([r[5] for r in rows], np.int32),
        obs=np.stack([r[6] for r in rows]))
    print("obs_general cases", len(rows))


def make_td():
    import torch
    g = torch.Generator().manual_seed(0)
    r = torch.tensor(REWARD_LUT, dtype=torch.float32)[torch.randint(0, 5, (2048, 1), generator=g)]
    v = torch.randn(2048, 1, generator=g); vn = torch.randn(2048, 1, generator=g)
    gamma = 0.99
    target_v = r + gamma * vn  # soa/agent/PPO.py:113
    adv = target_v - v         # soa/agent/PPO.py:114
    np.savez_compressed(os.path.join(HERE, "td_adv.npz"), r=r.numpy(), v=v.numpy(), v_next=vn.numpy(),
                        gamma=np.float32(gamma), adv=adv.numpy(), target_v=target_v.numpy())


if __name__ == "__main__":
    make_traj(4)
    make_traj(6)
    make_kat()
    make_obs_general()
    make_td()
