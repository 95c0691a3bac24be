"""CPU: the PPO host layer (twoarmy_b200.ppo) against fixtures produced by the REFERENCE's own
networks and PPO.update (tests/golden/make_golden_ppo.py -> ppo_ref.npz), plus the 2-rank gloo
test of the gradient all-reduce path."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _ppo():
    import twoarmy_b200
    return importlib.import_module(twoarmy_b200.__name__ + ".ppo")


def _sums(net):
    return np.array([p.detach().double().sum().item() for p in net.parameters()])


def test_networks_match_reference_under_same_seed(golden):
    """Same construction / init order as all_net.py: under torch.manual_seed(0) every parameter
    tensor and the outputs equal the reference's."""
    P = _ppo()
    fx = golden("ppo_ref.npz")
    torch.set_num_threads(1)
    torch.manual_seed(0)
    actor, critic = P.Net_PPO_actor(), P.Net_PPO_critic()
    np.testing.assert_allclose(_sums(actor), fx["actor_sums"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(_sums(critic), fx["critic_sums"], rtol=0, atol=1e-9)
    s, p, g = (torch.from_numpy(fx[k]) for k in ("x_s", "x_p", "x_g"))
    with torch.no_grad():
        prob = actor(s, p, g).numpy()
        val = critic(s, p, g).numpy()
    np.testing.assert_allclose(prob, fx["actor_prob"], rtol=1e-6, atol=1e-7)   # fp32, tolerance 1e-6 relative
    np.testing.assert_allclose(val, fx["critic_v"], rtol=1e-5, atol=1e-6)
    # parameter names are the reference's (checkpoints interchange)
    names = [n for n, _ in actor.named_parameters()]
    assert names[0] == "bone1.cnn_base.0.weight" and names[-1] == "A.bias" and "bone1.positionnet.weight" in names


def test_update_matches_reference(golden):
    """PPO.update (PPO.py:103-158) on the reference's record layout: same sampler stream, same
    losses, same Adam steps -> same parameters (fp32, 1e-5 relative on the per-tensor sums)."""
    P = _ppo()
    fx = golden("ppo_ref.npz")
    torch.set_num_threads(1)
    torch.manual_seed(0)
    agent = P.PPO(device="cpu", autocast=False)
    agent.K_epochs, agent.batch_size = 2, 32
    buf = {k: torch.from_numpy(fx[f"buf_{k}"]) for k in ("s", "a", "p", "g", "r", "a_logp")}
    torch.manual_seed(1)
    agent.update(buf)
    np.testing.assert_allclose(_sums(agent.actor), fx["upd_actor_sums"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(_sums(agent.critic), fx["upd_critic_sums"], rtol=1e-5, atol=1e-5)
    assert agent.update_count == 2 * 3


def test_codes_decode_to_matrix_env_values():
    P = _ppo()
    codes = torch.tensor([0, 1, 2, 4], dtype=torch.uint8)
    assert P.decode_matrix(codes).tolist() == pytest.approx([0.9, -0.9, -0.5, 0.3])


def test_select_action_is_batched_and_uses_newest_four_frames():
    P = _ppo()
    torch.manual_seed(0)
    agent = P.PPO(device="cpu", autocast=False)
    N = 5
    s = torch.randint(0, 3, (N, 5, 289), dtype=torch.uint8)
    p = torch.randint(1, 16, (N, 5, 2)).float()
    g = torch.tensor([[2.0, 14.0]]).repeat(N, 1)
    torch.manual_seed(5)
    a, lp = agent.select_action(s, p, g)
    assert a.shape == (N,) and lp.shape == (N,) and a.dtype == torch.int64
    s2 = s.clone()
    s2[:, 0] = 1  # the oldest frame is not an input of the policy (PPO.py:74-77)
    torch.manual_seed(5)
    a2, lp2 = agent.select_action(s2, p, g)
    assert torch.equal(a, a2) and torch.equal(lp, lp2)
    with torch.no_grad():
        prob = agent.actor(P.decode_matrix(s[:, 1:5]), p[:, 1:5], g)
    assert torch.allclose(lp, torch.log(prob[torch.arange(N), a]), atol=1e-6)


def test_update_with_virtual_her_samples_equals_materialised_copies():
    """`src` indirection (her.relabel's virtual samples) == the same records physically copied with
    their g / r overridden, which is what her_func appends (env_buffer.py:118-139)."""
    P = _ppo()
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(4)
    n, m = 24, 10
    base = {"s": torch.randint(0, 3, (n, 5, 289), generator=g, dtype=torch.uint8), "p": torch.randint(1, 16, (n, 5, 2), generator=g).float(),
            "a": torch.randint(0, 5, (n, 1), generator=g), "r": torch.rand(n, 1, generator=g) - 0.5,
            "a_logp": torch.log(torch.rand(n, 1, generator=g) * 0.3 + 0.1)}
    src = torch.cat([torch.arange(n), torch.randint(0, n, (m,), generator=g)])
    gg = torch.cat([torch.tensor([[2.0, 14.0]]).repeat(n, 1), torch.randint(1, 16, (m, 2), generator=g).float()])
    rr = torch.cat([base["r"], torch.full((m, 1), 0.9)])
    out = []
    for mode in ("virtual", "copied"):
        torch.manual_seed(0)
        agent = P.PPO(device="cpu", autocast=False)
        agent.K_epochs = 1
        if mode == "virtual":
            buf = dict(base, src=src, g=gg, r=rr)
        else:
            buf = {k: v[src] for k, v in base.items()}
            buf.update(g=gg, r=rr)
        torch.manual_seed(9)
        agent.update(buf, minibatch=17)
        out.append(_sums(agent.actor).tolist() + _sums(agent.critic).tolist())
    np.testing.assert_allclose(out[0], out[1], rtol=0, atol=0)


def test_folded_conv1_is_exact():
    """Upsample(4)+Conv2d(4,64,4,2)+ReLU == the 2x2-patch GEMM with phase-summed taps (fp32 round-off)."""
    P = _ppo()
    torch.manual_seed(0)
    net = P.TINet()
    x = torch.randn(3, 4, 17, 17)
    with torch.no_grad():
        want = torch.relu(net.cnn_base[0](net.upsamplingnearest(x)))
        got = net._conv1_folded(x)
    assert got.shape == want.shape and float((got - want).abs().max()) < 1e-5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ddp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    P = _ppo()
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    agent = P.PPO(device="cpu", autocast=False)
    agent.broadcast_parameters()
    agent.K_epochs = 1
    g = torch.Generator().manual_seed(11)
    n = 32
    full = {"s": torch.randint(0, 3, (n, 5, 289), generator=g, dtype=torch.uint8), "p": torch.randint(1, 16, (n, 5, 2), generator=g).float(),
            "a": torch.randint(0, 5, (n, 1), generator=g), "g": torch.tensor([[2.0, 14.0]]).repeat(n, 1),
            "r": torch.rand(n, 1, generator=g) - 0.5, "a_logp": torch.log(torch.rand(n, 1, generator=g) * 0.3 + 0.1)}
    half = n // world
    shard = {k: v[rank * half:(rank + 1) * half] for k, v in full.items()}
    agent.update(shard, minibatch=half)  # one optimiser step; gradients averaged over ranks
    ret[rank] = _sums(agent.actor).tolist() + _sums(agent.critic).tolist()
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce_equals_single_rank_full_batch():
    """world_size 2 (gloo, CPU): each rank updates on half the batch with the gradient
    all-reduce; the result equals one rank updating on the whole batch, and both ranks agree."""
    import torch.multiprocessing as mp
    P = _ppo()
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ddp_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] == pytest.approx(ret[1], rel=0, abs=0)
    # single-rank reference on the full batch
    torch.set_num_threads(1)
    torch.manual_seed(0)
    agent = P.PPO(device="cpu", autocast=False)
    agent.K_epochs = 1
    g = torch.Generator().manual_seed(11)
    n = 32
    full = {"s": torch.randint(0, 3, (n, 5, 289), generator=g, dtype=torch.uint8), "p": torch.randint(1, 16, (n, 5, 2), generator=g).float(),
            "a": torch.randint(0, 5, (n, 1), generator=g), "g": torch.tensor([[2.0, 14.0]]).repeat(n, 1),
            "r": torch.rand(n, 1, generator=g) - 0.5, "a_logp": torch.log(torch.rand(n, 1, generator=g) * 0.3 + 0.1)}
    agent.update(full, minibatch=n)
    want = _sums(agent.actor).tolist() + _sums(agent.critic).tolist()
    np.testing.assert_allclose(np.array(ret[0]), np.array(want), rtol=1e-5, atol=1e-5)


def _ddp_unequal_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    P = _ppo()
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    agent = P.PPO(device="cpu", autocast=False)
    agent.broadcast_parameters()
    agent.K_epochs = 2
    g = torch.Generator().manual_seed(11 + rank)
    n = 40 if rank == 0 else 23           # hindsight relabels make B data-dependent per rank
    buf = {"s": torch.randint(0, 3, (n, 5, 289), generator=g, dtype=torch.uint8), "p": torch.randint(1, 16, (n, 5, 2), generator=g).float(),
           "a": torch.randint(0, 5, (n, 1), generator=g), "g": torch.tensor([[2.0, 14.0]]).repeat(n, 1),
           "r": torch.rand(n, 1, generator=g) - 0.5, "a_logp": torch.log(torch.rand(n, 1, generator=g) * 0.3 + 0.1)}
    agent.update(buf, minibatch=8)        # rank 0 alone would run 5 steps per epoch, rank 1 three: both must run 2 (23 // 8)
    ret[rank] = (agent.update_count, _sums(agent.actor).tolist() + _sums(agent.critic).tolist())
    dist.destroy_process_group()


def test_two_rank_gloo_unequal_batch_sizes_agree_on_the_step_count():
    """with --her the number of samples differs per rank; every rank must issue the same number of gradient
    all-reduces (a mismatch deadlocks NCCL).  Both ranks run min(B) // bs full minibatches per epoch and end with
    identical parameters."""
    import torch.multiprocessing as mp
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ddp_unequal_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0][0] == ret[1][0] == 2 * (23 // 8)
    assert ret[0][1] == pytest.approx(ret[1][1], rel=0, abs=0)


def test_merged_parity_planes_are_the_strided_data_gradient():
    """conv1.parity_class_weights (torch path): ONE stride-1 convolution of dz with the merged [4*cin, cout, 2, 2] weight
    yields, per output-channel block pa*2+pb, the data gradient of the k x k stride-2 convolution at input pixels
    (2i+pa, 2j+pb) -- i.e. conv_transpose2d, exactly (k = 3 and k = 4: TINet's second / third convolution)."""
    import importlib
    import torch.nn.functional as F
    import twoarmy_b200 as pkg
    C1 = importlib.import_module(pkg.__name__ + ".conv1")
    g = torch.Generator().manual_seed(0)
    for cout, cin, k, oh in [(5, 3, 3, 6), (4, 6, 4, 7), (8, 8, 3, 16)]:
        w = torch.randn(cout, cin, k, k, generator=g, dtype=torch.float64)
        dz = torch.randn(2, cout, oh, oh, generator=g, dtype=torch.float64)
        planes = F.conv2d(dz, C1.parity_class_weights(w), padding=1)
        dx = F.conv_transpose2d(dz, w, stride=2)
        assert planes.shape == (2, 4 * cin, oh + 1, oh + 1)
        for c in range(4):
            want = dx[:, :, (c >> 1)::2, (c & 1)::2]
            got = planes[:, c * cin:(c + 1) * cin]
            assert torch.allclose(got[:, :, :want.shape[2], :want.shape[3]], want, rtol=0, atol=1e-12)
            assert float(got[:, :, want.shape[2]:].abs().max() if want.shape[2] < oh + 1 else 0.0) == 0.0   # the extra row is zero


def test_next_value_hint_reuses_the_first_critic_pass():
    """RolloutBuffer.flat()'s `next_same` flags (record t + 1 continues record t's episode) let PPO.values read V(s') from the
    first critic pass; on a buffer whose records really chain that is the value of the reference's second pass
    (soa/agent/PPO.py:113-114), and an update() built on it has the same losses."""
    P = _ppo()
    torch.manual_seed(0)
    agent = P.PPO(device="cpu", autocast=False)
    T, N = 6, 4
    buf = P.RolloutBuffer(T, N, torch.device("cpu"))
    g = torch.Generator().manual_seed(3)
    frames = torch.tensor([0, 1, 2, 4], dtype=torch.uint8)[torch.randint(0, 4, (T + 4, N, 289), generator=g)]
    pos = torch.randint(1, 16, (T + 4, N, 2), generator=g).float()
    ended = torch.zeros((T, N), dtype=torch.uint8)
    ended[2, 1] = 1; ended[4, 3] = 1; ended[0, 0] = 1
    codes = torch.tensor([0, 1, 2, 4], dtype=torch.uint8)
    s_prev = p_prev = None
    for t in range(T):
        # the rollout's rule (VecRollout / ta_stack_push): record t = record t - 1 shifted by one frame + the new frame, or, where
        # the previous step ended the episode, four fresh frames + the new frame
        if t == 0:
            s_t = torch.stack([frames[j] for j in range(5)], 1)
            p_t = torch.stack([pos[j] for j in range(5)], 1)
        else:
            s_t = torch.cat([s_prev[:, 1:5], frames[t + 4][:, None]], 1)
            p_t = torch.cat([p_prev[:, 1:5], pos[t + 4][:, None]], 1)
            broke = ended[t - 1].bool()
            s_t[broke, :4] = codes[torch.randint(0, 4, (int(broke.sum()), 4, 289), generator=g)]
            p_t[broke, :4] = torch.randint(1, 16, (int(broke.sum()), 4, 2), generator=g).float()
        buf.store(s_t, torch.randint(0, 5, (N,), generator=g), p_t, torch.rand(N, generator=g) - 0.5, ended[t].float(),
                  torch.log(torch.rand(N, generator=g) * 0.5 + 0.1))
        buf.ended[t].copy_(ended[t])
        s_prev, p_prev = s_t, p_t
    buf.g.copy_(torch.tensor([[2.0, 14.0]]).repeat(N, 1))
    flat = buf.flat()
    nxt = flat["next_same"].view(T, N)
    assert not bool(nxt[T - 1].any()) and bool((nxt[:T - 1] == (ended[:T - 1] == 0)).all()) and flat["next_stride"] == N
    with torch.no_grad():
        v2, vn2 = agent.values(flat["s"], flat["p"], flat["g"])
        v1, vn1 = agent.values(flat["s"], flat["p"], flat["g"], next_same=flat["next_same"], next_stride=N)
    assert torch.equal(v1, v2)
    assert torch.allclose(vn1, vn2, atol=1e-6)
    losses = []
    for share in (True, False):
        torch.manual_seed(0)
        ag = P.PPO(device="cpu", autocast=False)
        ag.share_next_value, ag.K_epochs, ag.batch_size = share, 2, 8
        torch.manual_seed(1)
        losses.append(ag.update(flat, sampler_generator=torch.default_generator))
    np.testing.assert_allclose(losses[0], losses[1], rtol=1e-5, atol=1e-7)
