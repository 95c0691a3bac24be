"""Scripted trajectories of the UNMODIFIED reference that force the rare branches of Twoarmy_v4.step
(SURVEY.md section 4's list): every patrol-collision direction (obstacles1 moving up / down onto the agent,
obstacles2 moving left / right onto it), each adjacency penalty (below / left / right of obstacles2, left of
obstacles1), the risk_count > 5 truncation, the room-2 bonus, wall blocks in the way, step-50 truncation and the
action >= 7 clamp.  The random / goal-seeking policies of make_golden.py reach them only a handful of times
(4 + 1 collisions in 40 x 160 steps).

Run only where the reference exists:  python tests/golden/make_golden_scripted.py
Writes tests/golden/traj_scripted_v4.npz in exactly the format of traj_v4.npz (checked by tests/traj_check.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (installs the import shim, imports the reference)

# (target cell to reach and hold, what it provokes)
TARGETS = [
    ((12, 3), "obstacles1 moving UP onto the agent"),
    ((12, 7), "obstacles1 moving DOWN onto the agent"),
    ((5, 4), "obstacles2 moving LEFT onto the agent"),
    ((11, 5), "obstacles2 moving RIGHT onto the agent"),
    ((11, 5), "left of obstacles1 (x == 11): adjacency penalty, risk_count"),
    ((8, 6), "below obstacles2: adjacency penalty, risk_count > 5"),
    ((8, 9), "below the mid balls: risk_count > 5"),
    ((6, 8), "the gap cell a mid ball moves onto"),
    ((10, 8), "the other gap cell"),
    ((4, 4), "left of obstacles2's leftmost position"),
    ((11, 4), "right of obstacles2 / hit by its right column"),
    ((7, 4), "inside obstacles2's row, left part: hit by either direction"),
    ((9, 5), "inside obstacles2's row, right part"),
    ((6, 5), "inside obstacles2's row, leftmost free cell"),
    ((10, 4), "inside obstacles2's row, right end"),
    ((3, 15), "never moves: step-50 truncation; clamped actions"),
]


def scripted_action(env, target, rng, t, lazy):
    if lazy:  # stay put; every 7th action is >= 7 (clamped to 0 = left, twoarmy_v4.py:84-85)
        return int(rng.choice([7, 11, 200])) if t % 7 == 3 else 6
    a = MG.bfs_action(env, target, rng)
    if a is None:   # blocked (wall block / ball in the way): wait
        return 6
    return a


def run(version=4, T=220, seed0=424200):
    N = len(TARGETS)
    env_id = f"MiniGrid-twoarmy-17x17-v{version}"
    gym = MG.gym
    view_sets = {}
    for view in (17, 7):
        out = dict(
            actions=np.zeros((T, N), np.int32), forced_reset=np.zeros((T, N), np.uint8),
            draws=np.full((T, N, 8), 0xFF, np.uint8), reward_idx=np.zeros((T, N), np.uint8),
            reward=np.zeros((T, N), np.float64), term=np.zeros((T, N), np.uint8),
            trunc=np.zeros((T, N), np.uint8), agent=np.zeros((T, N, 2), np.int8),
            grid=np.zeros((T, N, 289), np.uint8), flags=np.zeros((T, N, 10), np.int32),
            balls=np.zeros((T, N, 10, 2), np.int8), obs=np.zeros((T, N, view, view, 3), np.uint8),
            reset_obs=np.zeros((N, view, view, 3), np.uint8),
            post_reset_obs=np.zeros((T, N, view, view, 3), np.uint8),
        )
        stats = []
        for n, (target, what) in enumerate(TARGETS):
            np.random.seed(seed0 + n)
            rng = np.random.default_rng(77 + n)
            env = gym.make(env_id, agent_view_size=view)
            out["reset_obs"][n] = env.reset()["image"]
            hits = dict(o1_up=0, o1_down=0, o2_left=0, o2_right=0, mid=0, risk_trunc=0, room2=0, t50=0, adj=0)
            with MG.DrawRecorder(version) as rec:
                for t in range(T):
                    a = scripted_action(env, target, rng, t, lazy=(target == (3, 15)))
                    out["actions"][t, n] = a
                    pre = dict(up1=env.up1, right2=env.right2, risk=env.risk_count, steps=env.step_count,
                               o1=[o.cur_pos for o in env.obstacles1], o2=[o.cur_pos for o in env.obstacles2])
                    rec.begin_step()
                    obs, r, te, tr, _ = env.step(a)
                    out["draws"][t, n] = rec.end_step()
                    out["reward_idx"][t, n] = MG.REWARD_LUT.index(r)
                    out["reward"][t, n] = r
                    out["term"][t, n], out["trunc"][t, n] = te, tr
                    out["agent"][t, n] = env.agent_pos
                    out["grid"][t, n] = MG.grid_codes(env)
                    out["flags"][t, n] = MG.flags(env)
                    out["balls"][t, n] = MG.ball_positions(env)
                    out["obs"][t, n] = obs["image"]
                    ap = tuple(env.agent_pos)
                    if r == -0.9:
                        o1 = [tuple(o.cur_pos) for o in env.obstacles1 if o.cur_pos is not None]
                        o2 = [tuple(o.cur_pos) for o in env.obstacles2 if o.cur_pos is not None]
                        if ap in o1:
                            hits["o1_up" if pre["up1"] else "o1_down"] += 1
                        elif ap in o2:
                            hits["o2_right" if pre["right2"] else "o2_left"] += 1
                        else:
                            hits["mid"] += 1
                    if r == -0.1:
                        hits["adj"] += 1
                        if tr and pre["risk"] >= 5:
                            hits["risk_trunc"] += 1
                    if r == 0.2:
                        hits["room2"] += 1
                    if tr and r == -0.01 and pre["steps"] == 49:
                        hits["t50"] += 1
                    if te or tr:
                        out["post_reset_obs"][t, n] = env.reset()["image"]
            stats.append((what, hits))
        view_sets[view] = (out, stats)
    a, stats = view_sets[17]
    b, _ = view_sets[7]
    for k in a:
        if k not in ("obs", "reset_obs", "post_reset_obs"):
            assert np.array_equal(a[k], b[k]), k
    a["obs7"], a["reset_obs7"], a["post_reset_obs7"] = b["obs"], b["reset_obs"], b["post_reset_obs"]
    a["obs17"], a["reset_obs17"], a["post_reset_obs17"] = a.pop("obs"), a.pop("reset_obs"), a.pop("post_reset_obs")
    total = {}
    for what, h in stats:
        print(f"{what:60s} {h}")
        for k, v in h.items():
            total[k] = total.get(k, 0) + v
    print("TOTAL", total)
    a["coverage_names"] = np.array(sorted(total))
    a["coverage_counts"] = np.array([total[k] for k in sorted(total)], np.int32)
    np.savez_compressed(os.path.join(HERE, f"traj_scripted_v{version}.npz"), **a)
    return total


if __name__ == "__main__":
    run()
