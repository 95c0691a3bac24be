"""GPU: the single-env gym facade is a drop-in for gym.make(id) -- with the global numpy stream
seeded as in soa/train_ppo.py:51 it reproduces the digests the REFERENCE produced (SURVEY.md
section 4), i.e. identical obs / reward / flags / grid / agent for 5000 steps and 100 resets."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

SURVEY_DIGESTS = {(4, 17): "6dfb44abb94dd12a", (4, 7): "37f72f56ae52bacc",
                  (6, 17): "3015ddbb3f89096f", (6, 7): "3c292dc5e32c818c"}


def _gym_env():
    import importlib
    import twoarmy_b200
    return importlib.import_module(twoarmy_b200.__name__ + ".gym_env")


@pytest.mark.parametrize("version,view", list(SURVEY_DIGESTS))
def test_facade_reproduces_reference_digests(version, view):
    G = _gym_env()
    np.random.seed(9981)
    env = G.make(f"MiniGrid-twoarmy-17x17-v{version}", agent_view_size=view, seed=9981, new_step_api=True)
    acts = np.random.RandomState(7).choice(np.array([0, 1, 2, 3, 6]), size=5000)
    h = hashlib.sha256()
    h.update(env.reset()["image"].tobytes())
    total, eps = 0.0, 0
    for a in acts:
        obs, r, te, tr, info = env.step(int(a))
        assert isinstance(r, float) and isinstance(te, bool) and isinstance(tr, bool) and info == {}
        h.update(obs["image"].tobytes()); h.update(np.float64(r).tobytes()); h.update(bytes([te, tr]))
        h.update(env.grid.encode().tobytes()); h.update(bytes(env.agent_pos))
        total += r
        if te or tr:
            eps += 1
            h.update(env.reset()["image"].tobytes())
    assert h.hexdigest()[:16] == SURVEY_DIGESTS[(version, view)]
    assert eps == 100 and round(total, 2) == (-50.99 if version == 4 else -50.27)


def test_facade_surface_matches_reference_boundary():
    G = _gym_env()
    env = G.make("MiniGrid-twoarmy-17x17-v6")
    obs = env.reset()
    assert set(obs) == {"image", "direction", "mission"} and obs["direction"] == 3
    assert obs["image"].shape == (17, 17, 3) and obs["image"].dtype == np.uint8
    assert obs["mission"] == env.mission == "get to the green goal square"
    assert env.grid.height == 17 and len(env.grid.grid) == 289
    assert env.grid.grid[0].type == "wall" and env.grid.grid[15 * 17 + 3] is None
    assert env.grid.grid[2 * 17 + 14].type == "goal" and env.grid.grid[8 * 17 + 7].type == "ball"
    assert env.agent_pos == (3, 15) and env.goal_pos == (14, 2) and env.max_steps == 50
    assert [o.cur_pos for o in env.obstacles] == [(7, 8), (8, 8), (9, 8)]
    assert all(o.cur_pos is None for o in env.obstacles1 + env.obstacles2)
    assert int(env.actions.left) == 0 and int(env.actions.done) == 6 and env.action_space.n == 7
    assert env.get_full_render().shape == (17 * 32, 17 * 32, 3)
    with pytest.raises(AttributeError):
        env.step(4)  # minigrid.py:1397
    with pytest.raises(KeyError):
        G.make("MiniGrid-twoarmy-17x17-v0")
    # scripted known answer: goal at step 24
    env.reset()
    rew = [env.step(a)[1] for a in [2] * 6 + [1] * 7 + [2] * 7 + [1] * 4]
    assert rew == [-0.01] * 8 + [-0.1] * 5 + [-0.01, 0.2] + [-0.01] * 8 + [0.9]
    assert env.agent_pos == (14, 2)
