"""Generates render_ref.npz by RUNNING THE REFERENCE's get_full_render (gym_minigrid/minigrid.py:1514-1563,
Grid.render :712-747, render_tile :662-710, rendering.py).  Run only where the reference exists.

For tile sizes 8 and 17, highlight off / on and agent_view_size 17 / 7: a v4 env is stepped with random
actions; at a handful of steps the fixture keeps the state (grid codes y*17+x, agent x,y) and the RGB frame."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import gym  # noqa: E402

CODE = {None: 0, "wall": 1, "ball": 2, "goal": 3}


def main():
    out = {}
    k = 0
    for ts, hl, view in ((8, False, 17), (8, True, 17), (17, False, 17), (8, True, 7), (5, True, 7)):
        np.random.seed(31 + k)
        env = gym.make("MiniGrid-twoarmy-17x17-v4", tile_size=ts, highlight=hl, agent_view_size=view)
        env.reset()
        acts = np.random.RandomState(3 + k).choice(np.array([0, 1, 2, 2, 2, 3, 6]), size=40)
        for t, a in enumerate(acts):
            _, _, te, tr, _ = env.step(int(a))
            if t in (0, 9, 21, 39):
                grid = np.array([CODE[None if c is None else c.type] for c in env.grid.grid], np.uint8)
                out[f"r{k}_meta"] = np.array([ts, int(hl), view, env.agent_pos[0], env.agent_pos[1]], np.int32)
                out[f"r{k}_grid"] = grid
                out[f"r{k}_img"] = env.get_full_render()
                k += 1
            if te or tr:
                env.reset()
    np.savez_compressed(os.path.join(HERE, "render_ref.npz"), **out)
    print("wrote render_ref.npz with", k, "frames", {out[f"r{i}_img"].shape for i in range(k)})


if __name__ == "__main__":
    main()
