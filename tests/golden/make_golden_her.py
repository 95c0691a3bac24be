"""Generates her_ref.npz by RUNNING THE REFERENCE's Buffer_gridworld.her_func
(/root/reference/soa/env_buffer.py:101-143) on synthetic episodes.  Run only where the reference
exists:   python tests/golden/make_golden_her.py

For each case: an episode of L records is stored from `start` in a 2048-slot buffer (the reference's
record dtype, soa/train_ppo.py:93-97), np.random.seed(seed) is set, her_func() runs, and the fixture
keeps the episode (p, g, r, d), the seed, and everything the call appended: counter, full flag and
the appended records' (p4, g, r, d) plus which source record each one copies."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import env_buffer  # noqa: E402


def walk(rng, L):
    """A plausible agent track: (y, x) floats with revisits (the agent often bumps into walls)."""
    y, x = 15, 3
    out = []
    for _ in range(L):
        dy, dx = [(0, -1), (0, 1), (-1, 0), (1, 0), (0, 0)][rng.randint(5)]
        if rng.rand() < 0.35:
            dy = dx = 0
        y, x = min(15, max(1, y + dy)), min(15, max(1, x + dx))
        out.append((y, x))
    return np.array(out, np.float32)


def main():
    rng = np.random.RandomState(11)
    tr = np.dtype([("s", np.float32, (5, 289)), ("a", np.int64, (1,)), ("p", np.float32, (5, 2)), ("g", np.float32, (2,)),
                   ("r", np.float32, (1,)), ("d", np.float32, (1,)), ("a_logp", np.float32, (1,))])
    out = {}
    cases = [(50, 0), (50, 300), (1, 10), (2, 10), (3, 0), (17, 1990), (50, 1900), (50, 2040), (9, 77), (50, 1000), (33, 500), (50, 1500), (50, 1960), (40, 1975)]
    for ci, (L, start) in enumerate(cases):
        b = env_buffer.Buffer_gridworld()
        b.grid_size = 17
        b.transition = tr
        b.buffer_capacity = 2048
        b.buffer = np.zeros(2048, dtype=tr)
        p4 = walk(rng, L)
        ep = np.zeros(L, dtype=tr)
        ep["p"][:, 4] = p4
        ep["p"][:, 0:4] = rng.randint(1, 16, size=(L, 4, 2))
        ep["s"][:, 0, 0] = np.arange(L) + 1000 * ci  # tags the source record
        ep["a"][:, 0] = rng.randint(0, 5, size=L)
        ep["g"] = np.array([2.0, 14.0], np.float32)
        ep["r"][:, 0] = rng.choice(np.array([-0.01, -0.1, 0.2], np.float32), size=L)
        ep["a_logp"][:, 0] = -rng.rand(L)
        b.epo_counter_start = start
        b.counter = start
        for rec in ep:
            b.store(tuple(rec[k] for k in tr.names))
        # the reference's callers only run her_func right after the episode's last store
        cnt_before = b.counter
        seed = 100 + ci
        np.random.seed(seed)
        if b.counter == 0 and b.full:
            # the episode ended exactly on the ring boundary: epo_counter_end = -1 in the reference; skip
            continue
        b.her_func(max_steps=50, newgoal_size_in=4)
        n_new = (b.counter - cnt_before) % 2048
        idx = (cnt_before + np.arange(n_new)) % 2048
        new = b.buffer[idx]
        out[f"c{ci}_p4"] = p4
        out[f"c{ci}_r"] = ep["r"][:, 0]
        out[f"c{ci}_meta"] = np.array([L, start, seed, cnt_before, b.counter, int(b.full)], np.int64)
        out[f"c{ci}_new_src"] = (new["s"][:, 0, 0] - 1000 * ci).astype(np.int64)
        out[f"c{ci}_new_g"] = new["g"]
        out[f"c{ci}_new_r"] = new["r"][:, 0]
        out[f"c{ci}_new_d"] = new["d"][:, 0]
    np.savez_compressed(os.path.join(HERE, "her_ref.npz"), **out)
    print("wrote her_ref.npz:", sorted(k for k in out if k.endswith("_meta")))
    for k in sorted(out):
        if k.endswith("_meta"):
            print(k, out[k], "appended", len(out[k.replace("_meta", "_new_src")]))


if __name__ == "__main__":
    main()
