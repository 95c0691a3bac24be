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


def test_fold_decoder_tail_is_exact():
    """ConvTranspose2d(16, 1, 4, 2) + AvgPool2d(4) == the 3x3 stride-2 padding-1 convolution predictor.fold_decoder_tail
    builds (what ta_pred_decoder's last stage computes), in float64."""
    import torch.nn.functional as F
    M = _mod()
    torch.manual_seed(0)
    dec = M.Net_Decoder().double()
    x = torch.randn(3, 16, 33, 33, dtype=torch.float64)
    with torch.no_grad():
        want = dec.pool(dec.cnn_base[4](x))
        taps, b = M.fold_decoder_tail(dec.cnn_base[4].weight.detach(), dec.cnn_base[4].bias.detach())
        got = F.conv2d(x, taps.permute(2, 0, 1).unsqueeze(0), b.reshape(1), stride=2, padding=1)
    assert want.shape == got.shape == (3, 1, 17, 17)
    assert float((want - got).abs().max()) < 1e-12


def test_hand_scheduled_lstm_matches_nn_lstm():
    """LSTM._fast_forward (the GPU path's schedule: one input-gate GEMM per layer for the teacher-forced steps, a recurrent
    GEMM + cell update per step, [x | h] GEMMs for the self-fed steps) in fp32 on the CPU == the reference's forward
    through torch.nn.LSTM (all_net.py:76-98), with non-zero initial states."""
    M = _mod()
    torch.manual_seed(0)
    m = M.LSTM()
    m.h_0, m.c_0 = torch.randn(3, 1024) * 0.1, torch.randn(3, 1024) * 0.1
    z = torch.randn(5, 4, 64, 4, 4)
    with torch.no_grad():
        want, zc = m(z)
        got = m._fast_forward(z.reshape(5, 4, 1024), torch.float32).reshape(5, 7, 64, 4, 4)
    assert want.shape == got.shape and zc.shape == (5, 4, 1024)
    assert float((want - got).abs().max()) < 1e-5 * max(1.0, float(want.abs().max()))


def test_stack_arrays_fold_eval_batchnorm_and_layouts():
    """ppo_predictor._stack_arrays (what ta_pred_encoder / ta_pred_decoder read): convolution taps in [ky][kx][ci][co]
    order and eval-mode BatchNorm + bias folded into scale / shift reproduce the modules layer by layer (float32, CPU)."""
    import torch.nn.functional as F
    M = _mod()
    torch.manual_seed(1)
    agent = M.ppo_predictor(device="cpu", autocast=False)
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for m in agent.encoder.cnn_base:
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.2)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.weight.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    agent.encoder.eval(); agent.decoder.eval()
    e, d = agent._stack_arrays()
    enc, dec = agent.encoder.cnn_base, agent.decoder.cnn_base
    with torch.no_grad():
        x = torch.randn(2, 1, 68, 68, generator=g)
        for (ci, w_key, s_key, t_key, stride) in ((0, "w1", "s1", "t1", 2), (3, "w2", "s2", "t2", 4), (6, "w3", "s3", "t3", 2)):
            conv, bn = enc[ci], enc[ci + 1]
            want = torch.relu(bn(conv(x)))
            w = e[w_key].reshape(conv.weight.shape) if w_key == "w1" else e[w_key].permute(3, 2, 0, 1)   # back to [co][ci][ky][kx]
            got = torch.relu(F.conv2d(x, w, None, stride=stride) * e[s_key].view(1, -1, 1, 1) + e[t_key].view(1, -1, 1, 1))
            assert float((want - got).abs().max()) < 1e-4 * max(1.0, float(want.abs().max())), w_key
            x = want
        z = torch.randn(2, 64, 4, 4, generator=g)
        for (ci, w_key, b_key, stride) in ((0, "w1", "b1", 2), (2, "w2", "b2", 4)):
            want = torch.relu(dec[ci](z))
            got = torch.relu(F.conv_transpose2d(z, d[w_key].permute(2, 3, 0, 1), d[b_key], stride=stride))      # [ci][co][ky][kx]
            assert float((want - got).abs().max()) < 1e-5 * max(1.0, float(want.abs().max())), w_key
            z = want
        want = agent.decoder.pool(dec[4](z))
        got = F.conv2d(z, d["w3"].permute(2, 0, 1).unsqueeze(0), torch.tensor([d["b3"]]), stride=2, padding=1)
        assert float((want - got).abs().max()) < 1e-5 * max(1.0, float(want.abs().max()))
