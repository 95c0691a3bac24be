"""Generates pre_her_ref.npz by RUNNING THE REFERENCE's Buffer_gridworld.pre_store / pre_her_func
(/root/reference/soa/env_buffer.py:90-99, 145-210) on synthetic episodes fed through the record-building
loop of soa/train_ppo_predictor.py:105-171 (9-frame sliding windows, records from the 5th step on, four
closing pads).  Run only where the reference exists:   python tests/golden/make_golden_pre_her.py

Every frame is tagged with the step that produced it (s[f, 0] = step, -1 = the reset frame), so the fixture
can say, for each record pre_her_func appended, which step's 5-frame record sits in frames 0..4 -- the part
PPO_Predictor.update reads (PPO_Predictor.py:124-163) -- next to its g, a[0], r[0], a_logp[0]."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import env_buffer  # noqa: E402
from make_golden_her import walk  # noqa: E402


def roll(stack, new):
    return np.append(np.delete(stack, 0, 0), [new], 0)     # train_ppo_predictor.py:126-139


def main():
    rng = np.random.RandomState(23)
    pt = np.dtype([("s", np.float64, (9, 289)), ("a", np.int64, (5, 1)), ("p", np.float64, (9, 2)), ("g", np.float64, (2,)),
                   ("r", np.float64, (5, 1)), ("d", np.int64, (5, 1)), ("a_logp", np.float64, (5, 1))])   # :103-106
    out = {}
    cases = [(50, 0), (50, 700), (30, 10), (12, 0), (6, 40), (5, 3), (4, 100), (2, 9), (50, 1200), (41, 300), (50, 1500), (23, 64)]
    for ci, (L, start) in enumerate(cases):
        b = env_buffer.Buffer_gridworld()
        b.grid_size = 17
        b.buffer_pre_capacity = 2048
        b.pre_transition = pt
        b.pre_buffer = np.zeros(2048, dtype=pt)
        b.pre_counter = start
        b.epo_counter_start = b.pre_counter                                   # :119
        pos = walk(rng, L).astype(np.float64)                                 # (y, x) after each step
        act = rng.randint(0, 5, size=L)
        rew = rng.choice(np.array([-0.01, -0.1, 0.2]), size=L)
        alp = -rng.rand(L)
        frame0 = np.zeros(289); frame0[0] = -1                                # the reset frame's tag
        s9 = np.tile(frame0, (9, 1)); p9 = np.tile(np.array([15.0, 3.0]), (9, 1))          # predata_reset, env_buffer.py:430-437
        a5, r5, d5, l5 = np.zeros((5, 1)), np.zeros((5, 1)), np.zeros((5, 1)), np.zeros((5, 1))   # :117
        goal = np.array([2.0, 14.0])

        def store():
            b.pre_store((np.array(s9, dtype="float32"), np.array(a5, dtype="int64"), np.array(p9, dtype="float32"),
                         np.array(goal, dtype="float32"), np.array(r5, dtype="float32"), np.array(d5, dtype="int64"),
                         np.array(l5, dtype="float32")))
        for t in range(L):
            frame = np.zeros(289); frame[0] = t
            done = int(t == L - 1)
            s9, p9 = roll(s9, frame), roll(p9, pos[t])
            a5, r5, d5, l5 = roll(a5, [act[t]]), roll(r5, [rew[t]]), roll(d5, [done]), roll(l5, [alp[t]])
            if t > 3:
                store()                                                       # :140-142
            if done:
                for _ in range(4):                                            # :145-160
                    s9, p9 = roll(s9, frame), roll(p9, pos[t])
                    a5, r5, d5, l5 = roll(a5, [act[t]]), roll(r5, [rew[t]]), roll(d5, [done]), roll(l5, [alp[t]])
                    store()
        cnt_before = b.pre_counter
        J = cnt_before - start
        seed = 300 + ci
        np.random.seed(seed)
        b.pre_her_func(max_steps=50, newgoal_size_in=4)
        n_new = b.pre_counter - cnt_before
        assert n_new >= 0 and not b.pre_full, "keep the cases away from the ring boundary"
        new = b.pre_buffer[cnt_before:cnt_before + n_new]
        out[f"c{ci}_pos"] = pos.astype(np.float32)
        out[f"c{ci}_p8"] = b.pre_buffer[start:cnt_before]["p"][:, 8].astype(np.float32)
        out[f"c{ci}_rew"] = rew.astype(np.float32)
        out[f"c{ci}_act"] = act.astype(np.int64)
        out[f"c{ci}_alp"] = alp.astype(np.float32)
        out[f"c{ci}_meta"] = np.array([L, start, seed, J, n_new], np.int64)
        out[f"c{ci}_new_tags"] = new["s"][:, :, 0].astype(np.int64)           # [n_new, 9] step of every frame
        out[f"c{ci}_new_g"] = new["g"].astype(np.float32)
        out[f"c{ci}_new_r0"] = new["r"][:, 0, 0].astype(np.float32)
        out[f"c{ci}_new_a0"] = new["a"][:, 0, 0].astype(np.int64)
        out[f"c{ci}_new_alp0"] = new["a_logp"][:, 0, 0].astype(np.float32)
        out[f"c{ci}_new_r"] = new["r"][:, :, 0].astype(np.float32)
        out[f"c{ci}_new_d"] = new["d"][:, :, 0].astype(np.int64)
        print(f"case {ci}: L={L} records={J} appended={n_new}")
    np.savez_compressed(os.path.join(HERE, "pre_her_ref.npz"), **out)
    print("wrote pre_her_ref.npz")


if __name__ == "__main__":
    main()
