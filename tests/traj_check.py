"""Shared trajectory comparator: drives an implementation (the CPU oracle or the CUDA vector
env) with a golden fixture's actions / recorded draws / forced resets and compares every
quantity the reference produced, bit for bit."""
import numpy as np

REWARD_LUT32 = np.array([-0.01, -0.1, -0.9, 0.2, 0.9], dtype=np.float64).astype(np.float32)
FLAG_NAMES = ["step_count", "step_move", "risk_count", "pone", "patrol", "up1", "right2", "upd_h",
              "upd_l", "first_room2"]


def check_traj(impl, fx, view, steps=None):
    """impl: object with .reset() -> obs[N,V,V,3]; .step(actions, draws) -> dict(obs, reward,
    terminated, truncated, consumed); .reset_masked(mask) -> None; .state() -> dict(grid, agent,
    flags, balls) in the fixture's conventions; .obs_now() -> current obs[N,V,V,3]."""
    T = fx["actions"].shape[0] if steps is None else steps
    sfx = str(view)
    obs0 = impl.reset()
    assert np.array_equal(obs0, fx["reset_obs" + sfx]), "reset obs"
    for t in range(T):
        out = impl.step(fx["actions"][t], fx["draws"][t])
        drawn = (fx["draws"][t][:, :7] != 0xFF)
        want_mask = (drawn * (1 << np.arange(7))).sum(1).astype(np.uint8)
        assert np.array_equal(out["consumed"], want_mask), f"t={t}: RNG call sites executed"
        assert np.array_equal(out["obs"], fx["obs" + sfx][t]), f"t={t}: obs"
        assert np.array_equal(out["reward"], REWARD_LUT32[fx["reward_idx"][t]]), f"t={t}: reward"
        assert np.array_equal(out["reward"], fx["reward"][t].astype(np.float32)), f"t={t}: reward f32"
        assert np.array_equal(out["terminated"].astype(np.uint8), fx["term"][t]), f"t={t}: terminated"
        assert np.array_equal(out["truncated"].astype(np.uint8), fx["trunc"][t]), f"t={t}: truncated"
        st = impl.state()  # state as the reference holds it right after env.step (before reset)
        assert np.array_equal(st["agent"], fx["agent"][t]), f"t={t}: agent_pos"
        assert np.array_equal(st["grid"], fx["grid"][t]), f"t={t}: grid"
        assert np.array_equal(st["balls"], fx["balls"][t]), f"t={t}: ball positions"
        for k, name in enumerate(FLAG_NAMES):
            assert np.array_equal(st["flags"][:, k], fx["flags"][t][:, k]), f"t={t}: {name}"
        done = (fx["term"][t] | fx["trunc"][t]).astype(bool)
        mask = done | fx["forced_reset"][t].astype(bool)
        if mask.any():
            impl.reset_masked(mask.astype(np.uint8))
            now = impl.obs_now()
            assert np.array_equal(now[mask], fx["post_reset_obs" + sfx][t][mask]), f"t={t}: post-reset obs"
