"""CPU: the PPO+predictor host layer (twoarmy_b200.predictor) against fixtures produced by the REFERENCE's
own networks and ppo_predictor.update (tests/golden/make_golden_predictor.py -> predictor_ref.npz), and the
9-frame record helper against a literal replay of train_ppo_predictor.py:105-171."""
import importlib

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _mod():
    import twoarmy_b200
    return importlib.import_module(twoarmy_b200.__name__ + ".predictor")


def _sums(net):
    return np.array([p.detach().double().sum().item() for p in net.parameters()])


def test_networks_and_forward_match_reference(golden):
    M = _mod()
    fx = golden("predictor_ref.npz")
    torch.set_num_threads(1)
    torch.manual_seed(0)
    agent = M.ppo_predictor(device="cpu", autocast=False)
    for name in ("actor", "critic", "encoder", "decoder", "predictor"):
        np.testing.assert_allclose(_sums(getattr(agent, name)), fx[f"{name}_sums"], rtol=0, atol=1e-9, err_msg=name)
    x, p, g = (torch.from_numpy(fx[k]) for k in ("x", "p", "g"))
    pred = agent.pred_states(x)[0]
    np.testing.assert_allclose(pred.numpy(), fx["pred"], rtol=1e-5, atol=1e-6)
    agent.actor.eval(); agent.critic.eval()
    with torch.no_grad():
        cat = agent._net_in(x)
        np.testing.assert_allclose(agent.actor(cat, p, g).numpy(), fx["actor_prob"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(agent.critic(cat, p, g).numpy(), fx["critic_v"], rtol=1e-5, atol=1e-6)
    assert [n for n, _ in agent.actor.named_parameters()][0] == "bone1.cnn_base.0.weight"
    assert agent.actor.bone1.cnn_base[0].weight.shape == (64, 8, 4, 4)


def test_update_matches_reference_on_its_nine_frame_records(golden):
    """ppo_predictor.update reads frames 0..4, a[:,0], r[:,0], a_logp[:,0] of the 9-frame records
    (PPO_Predictor.py:124-163): feeding exactly those to this repo's update gives the reference's
    parameters (fp32, 1e-5 on the per-tensor sums)."""
    M = _mod()
    fx = golden("predictor_ref.npz")
    torch.set_num_threads(1)
    torch.manual_seed(0)
    agent = M.ppo_predictor(device="cpu", autocast=False)
    agent.K_epochs, agent.batch_size = 1, 16
    buf = {"s": torch.from_numpy(fx["buf_s"][:, 0:5]).float(), "p": torch.from_numpy(fx["buf_p"][:, 0:5]).float(),
           "a": torch.from_numpy(fx["buf_a"][:, 0]), "g": torch.from_numpy(fx["buf_g"]).float(),
           "r": torch.from_numpy(fx["buf_r"][:, 0]).float(), "a_logp": torch.from_numpy(fx["buf_a_logp"][:, 0]).float()}
    torch.manual_seed(1)
    agent.update(buf)
    np.testing.assert_allclose(_sums(agent.actor), fx["upd_actor_sums"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(_sums(agent.critic), fx["upd_critic_sums"], rtol=1e-5, atol=1e-5)


def test_pre_transition_records_replay_the_reference_loop():
    """The (time index, env) table of pre_transition_records == the 9-frame stacks the reference builds
    with np.delete / np.append and stores with `if t > 3` plus four closing records
    (train_ppo_predictor.py:123-160), for episodes of assorted lengths."""
    M = _mod()
    rng = np.random.RandomState(0)
    T, N = 70, 5
    ended = np.zeros((T, N), bool)
    for e in range(N):
        t = -1
        while True:
            t += rng.randint(1, 14) if e else 50
            if t >= T:
                break
            ended[t, e] = True
    idx, env = M.pre_transition_records(torch.from_numpy(ended))
    frames = torch.arange(T * N, dtype=torch.float32).reshape(T, N, 1) + torch.zeros(1, 1, 3)  # frame id = t*N+e
    rec = M.gather_records(frames, torch.full((1, 3), -7.0), idx, env)
    assert rec.shape == (idx.shape[0], 9, 3)
    assert torch.equal(rec[..., 0], torch.where(idx < 0, torch.tensor(-7.0), (idx * N + env.unsqueeze(-1)).float()))
    got = {}
    for row, e in zip(idx.numpy(), env.numpy()):
        got.setdefault(int(e), []).append(tuple(int(v) for v in row))
    for e in range(N):
        want = []
        t = 0
        while t < T:
            start = t
            stack = [-1] * 9               # predata_reset tiles the frame reset() left (index -1)
            k = 0
            while t < T:
                stack = stack[1:] + [t]    # np.delete(x,0,0); np.append(x,[new],0)
                if k > 3:
                    want.append(tuple(stack))
                k += 1
                if ended[t, e]:
                    for _ in range(4):
                        stack = stack[1:] + [t]
                        want.append(tuple(stack))
                    t += 1
                    break
                t += 1
        assert sorted(got.get(e, [])) == sorted(want), e
