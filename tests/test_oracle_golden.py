"""CPU-only: pins the oracle (oracle/twoarmy_oracle.c) to the reference.

Every fixture under tests/golden/ was produced by the reference itself (make_golden.py); the
SURVEY.md section-4 digests were produced by the reference during the survey."""
import ctypes as C
import hashlib

import numpy as np
import pytest

from oracle import oracle as O
from traj_check import check_traj, FLAG_NAMES


class OracleImpl:
    def __init__(self, version, n, view):
        self.b = O.OracleBatch(version, n, view)

    def reset(self):
        return self.b.reset()

    def step(self, actions, draws):
        return self.b.step(actions, draws, autoreset=False)

    def reset_masked(self, mask):
        self.b.reset(mask)

    def obs_now(self):
        return self.b.obs()

    def state(self):
        e = self.b.envs
        balls = np.concatenate([e["mid"], e["o1"], e["o2"]], axis=1)
        flags = np.stack([e[k].astype(np.int32) for k in FLAG_NAMES], axis=1)
        return dict(grid=e["grid"], agent=np.stack([e["ax"], e["ay"]], 1).astype(np.int8), flags=flags,
                    balls=balls)


@pytest.mark.parametrize("version", [4, 6])
@pytest.mark.parametrize("view", [17, 7])
def test_oracle_matches_reference_trajectories(golden, version, view):
    fx = golden(f"traj_v{version}.npz")
    check_traj(OracleImpl(version, fx["actions"].shape[1], view), fx, view)


@pytest.mark.parametrize("view", [17, 7])
def test_oracle_matches_scripted_rare_branch_trajectories(golden, view):
    """tests/golden/traj_scripted_v4.npz (make_golden_scripted.py): reference trajectories steered into every
    patrol-collision direction, every adjacency penalty, risk_count > 5, the room-2 bonus, step-50 truncation
    and clamped actions -- and the fixture really contains each of them."""
    fx = golden("traj_scripted_v4.npz")
    cov = dict(zip(fx["coverage_names"].tolist(), fx["coverage_counts"].tolist()))
    for k in ("o1_up", "o1_down", "o2_left", "o2_right", "mid", "risk_trunc", "room2", "t50", "adj"):
        assert cov[k] >= 1, k
    assert int((fx["actions"] >= 7).sum()) >= 5
    check_traj(OracleImpl(4, fx["actions"].shape[1], view), fx, view)


SURVEY_DIGESTS = {(4, 17): "6dfb44abb94dd12a", (4, 7): "37f72f56ae52bacc",
                  (6, 17): "3015ddbb3f89096f", (6, 7): "3c292dc5e32c818c"}
SURVEY_DRAWS = {4: (2212, "67768b5ade107503"), 6: (200, "850e170d38b84205")}


@pytest.mark.parametrize("version,view", list(SURVEY_DIGESTS))
def test_oracle_reproduces_survey_digests(version, view):
    """np.random.seed(9981) + the reference's conditional np.random.choice sites, replayed by the
    oracle's MT19937 provider, must give the digest the reference gave."""
    e = O.OracleMT(version, view, 9981)
    acts = np.random.RandomState(7).choice(np.array([0, 1, 2, 3, 6]), size=5000)
    h = hashlib.sha256()
    h.update(e.reset().tobytes())
    total, eps, ndraw = 0.0, 0, 0
    hd = hashlib.sha256()
    for a in acts:
        obs, r, te, tr, cons, vals = e.step(int(a))
        for s in range(7):
            if cons >> s & 1:
                ndraw += 1
        h.update(obs.tobytes()); h.update(np.float64(r).tobytes()); h.update(bytes([te, tr]))
        h.update(e.encode_grid().tobytes())
        h.update(bytes([int(e.env["ax"][0]), int(e.env["ay"][0])]))
        total += r
        if te or tr:
            eps += 1
            h.update(e.reset().tobytes())
    assert h.hexdigest()[:16] == SURVEY_DIGESTS[(version, view)]
    assert eps == 100
    assert ndraw == SURVEY_DRAWS[version][0]
    assert round(total, 2) == (-50.99 if version == 4 else -50.27)


def test_numpy_legacy_choice_emulation():
    mt = (C.c_uint32 * 625)()
    O.lib().mt_seed(mt, C.c_uint32(9981))
    np.random.seed(9981)
    rs = np.random.RandomState(5)
    for _ in range(5000):
        lo, hi = [(0, 10), (9, 13), (6, 10), (4, 5), (0, 2)][rs.randint(5)]
        assert np.random.choice(range(lo, hi), 1).item() == O.lib().ora_np_choice(mt, lo, hi)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [([0] * 4, [0] * 2, [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kat:
        assert list(O.philox4x32_10(ctr, key)) == want


def test_scripted_goal_kat(golden):
    fx = golden("kat_v6_goal.npz")
    b = O.OracleBatch(6, 1, 17)
    b.reset()
    for t, a in enumerate(fx["actions"]):
        out = b.step([a], np.array([[255, 255, 255, 255, 255, 1, 1, 255]], np.uint8), autoreset=False)
        assert out["reward"][0] == np.float32(fx["reward"][t])
        assert out["terminated"][0] == fx["term"][t] and out["truncated"][0] == fx["trunc"][t]
        assert (int(b.envs["ax"][0]), int(b.envs["ay"][0])) == tuple(fx["agent"][t])
    assert out["terminated"][0] == 1 and t == 23


def test_gen_obs_general_and_window_agree(golden):
    fx = golden("obs_general.npz")
    for i in range(len(fx["ax"])):
        V = int(fx["view"][i])
        got = O.gen_obs_general(fx["grid"][i], int(fx["ax"][i]), int(fx["ay"][i]), int(fx["dir"][i]), V,
                                int(fx["stw"][i]))
        assert np.array_equal(got, fx["obs"][i][:V, :V]), i


def test_matrix_env_and_data_env(golden):
    for version in (4, 6):
        fx = golden(f"traj_v{version}.npz")
        impl = OracleImpl(version, fx["actions"].shape[1], 17)
        impl.reset()
        for t in range(60):
            impl.step(fx["actions"][t], fx["draws"][t])
            assert np.array_equal(impl.b.matrix().astype(np.float32), fx["matrix"][t])
            e = impl.b.envs
            assert np.array_equal(np.stack([e["ay"], e["ax"]], 1).astype(np.float32), fx["place"][t])
            mask = (fx["term"][t] | fx["trunc"][t] | fx["forced_reset"][t]).astype(np.uint8)
            if mask.any():
                impl.reset_masked(mask)


def test_td_advantage(golden):
    fx = golden("td_adv.npz")
    adv, tv = O.td_advantage(fx["r"], fx["v"], fx["v_next"], float(fx["gamma"]))
    assert np.array_equal(tv, fx["target_v"]) and np.array_equal(adv, fx["adv"])
    # GAE(lambda=0, no mask) is the same quantity (float64 vs fp32: 1e-5 relative)
    a64, r64 = O.gae(fx["r"].reshape(1, -1), fx["v"].reshape(1, -1), np.zeros((1, 2048), np.uint8),
                     0.99, 0.0, use_mask=False, v_next=fx["v_next"].reshape(1, -1))
    np.testing.assert_allclose(a64[0], fx["adv"][:, 0], rtol=1e-5, atol=1e-6)


def test_her_func_matches_reference(golden):
    """oracle.her_func (restating soa/env_buffer.py:101-143) against what the reference's own
    her_func appended to its buffer (tests/golden/her_ref.npz), same global numpy seed."""
    from oracle import oracle as O
    fx = golden("her_ref.npz")
    cases = sorted({k.split("_")[0] for k in fx})
    assert len(cases) >= 10
    n_nonempty = 0
    for c in cases:
        L, start, seed, cnt_before, counter, full = (int(v) for v in fx[f"{c}_meta"])
        if cnt_before != start + L:      # the episode itself wrapped the ring: the reference's slice is empty
            assert len(fx[f"{c}_new_src"]) == 0
            continue
        np.random.seed(seed)
        got = O.her_func(fx[f"{c}_p4"], fx[f"{c}_r"], start)
        np.testing.assert_array_equal(got["src"], fx[f"{c}_new_src"])
        np.testing.assert_array_equal(got["g"], fx[f"{c}_new_g"].reshape(-1, 2))
        np.testing.assert_array_equal(got["r"], fx[f"{c}_new_r"])
        np.testing.assert_array_equal(got["d"], fx[f"{c}_new_d"])
        assert got["counter"] % 2048 == counter % 2048 and got["full"] == bool(full)
        n_nonempty += len(got["src"]) > 0
    assert n_nonempty >= 6


def test_pre_her_func_matches_reference(golden):
    """oracle.pre_her_func and oracle.her_plan(first=4) (restating soa/env_buffer.py:145-210 over the episode's
    steps) against what the reference's own pre_her_func appended to its 9-frame buffer
    (tests/golden/pre_her_ref.npz, produced through the record loop of train_ppo_predictor.py:105-171): for every
    appended record, the step whose 5-frame record sits in frames 0..4, the goal and a[0] / r[0] / a_logp[0]."""
    from oracle import oracle as O
    fx = golden("pre_her_ref.npz")
    cases = sorted({k.split("_")[0] for k in fx})
    assert len(cases) >= 10
    n_nonempty = 0
    for c in cases:
        L, start, seed, J, n_new = (int(v) for v in fx[f"{c}_meta"])
        pos, rew, act, alp, tags = fx[f"{c}_pos"], fx[f"{c}_rew"], fx[f"{c}_act"], fx[f"{c}_alp"], fx[f"{c}_new_tags"]
        np.random.seed(seed)
        got = O.pre_her_func(fx[f"{c}_p8"], rew)
        assert len(got["step"]) == n_new
        if L >= 4:                       # record j <-> step j + 4; the 4 closing pads repeat the last position
            np.testing.assert_array_equal(fx[f"{c}_p8"], np.concatenate([pos[4:], np.repeat(pos[-1:], 4, 0)]))
        if n_new == 0:
            continue
        n_nonempty += 1
        step = got["step"]
        np.testing.assert_array_equal(tags[:, 4], step)
        for f in range(4):                # frames 0..3 = the four steps before (the reset frame before step 0)
            np.testing.assert_array_equal(tags[:, f], np.maximum(step - 4 + f, -1))
        np.testing.assert_array_equal(got["g"], fx[f"{c}_new_g"])
        np.testing.assert_array_equal(got["r0"], fx[f"{c}_new_r0"])
        np.testing.assert_array_equal(act[step], fx[f"{c}_new_a0"])
        np.testing.assert_array_equal(alp[step], fx[f"{c}_new_alp0"])
        # the vectorised form the device computes: same draws, plan -> (step, goal, r) per relabel slot
        np.random.seed(seed)
        done = np.zeros((L, 1), np.uint8); done[-1] = 1
        plan = O.her_plan(pos[:, 0:1], pos[:, 1:2], done, lambda ind, k, t1, e: np.random.choice(ind, size=k, replace=False), first=4)
        ps, pg, pr = [], [], []
        for slot in range(4):
            tt = np.nonzero(plan[:, 0, slot] != 0xFFFF)[0]
            v = plan[tt, 0, slot].astype(np.int64)
            ps += list(tt); pg += [((x >> 5) & 31, x & 31) for x in v]
            pr += [np.float32(0.9) if x & 0x8000 else rew[t] for t, x in zip(tt, v)]
        np.testing.assert_array_equal(np.array(ps), step)
        np.testing.assert_array_equal(np.array(pg, np.float32), got["g"])
        np.testing.assert_array_equal(np.array(pr, np.float32), got["r0"])
    assert n_nonempty >= 6


def test_render_tiles_compose_to_reference_frames(golden):
    """Host tile atlas (render.tile_atlas: my restatement of render_tile / rendering.py) composed per cell ==
    the reference's get_full_render, pixel for pixel, for tile sizes 5/8/17, highlight on/off, views 17/7."""
    import importlib
    import twoarmy_b200 as pkg
    R = importlib.import_module(pkg.__name__ + ".render")
    fx = golden("render_ref.npz")
    n = len([k for k in fx if k.endswith("_meta")])
    assert n >= 20
    for k in range(n):
        ts, hl, view, ax, ay = (int(v) for v in fx[f"r{k}_meta"])
        got = R.compose(fx[f"r{k}_grid"], (ax, ay), ts, bool(hl), view)
        assert np.array_equal(got, fx[f"r{k}_img"]), k
