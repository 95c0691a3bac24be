"""GPU: the PPO host layer on the device -- the GEMM form of the last conv equals the cuDNN conv,
bf16 autocast stays close to fp32, and one vectorised rollout + update runs end to end."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _ppo():
    import twoarmy_b200
    return importlib.import_module(twoarmy_b200.__name__ + ".ppo")


def test_gpu_forward_matches_cpu_forward_fp32():
    """Same parameters, fp32 on both sides: the GPU path (channels_last, conv4 as im2col+GEMM) equals
    the plain CPU path (the reference's layer sequence) to fp32 round-off (1e-4 absolute)."""
    P = _ppo()
    torch.manual_seed(0)
    actor, critic = P.Net_PPO_actor(), P.Net_PPO_critic()
    g = torch.Generator().manual_seed(1)
    lut = torch.tensor([0.9, -0.9, -0.5, 0.3])
    s = lut[torch.randint(0, 4, (64, 4, 289), generator=g)]
    p = torch.randint(1, 16, (64, 4, 2), generator=g).float()
    goal = torch.tensor([[2.0, 14.0]]).repeat(64, 1)
    with torch.no_grad():
        want_a, want_v = actor(s, p, goal), critic(s, p, goal)
        actor.cuda().to(memory_format=torch.channels_last); critic.cuda().to(memory_format=torch.channels_last)
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        got_a, got_v = actor(s.cuda(), p.cuda(), goal.cuda()).cpu(), critic(s.cuda(), p.cuda(), goal.cuda()).cpu()
    assert torch.allclose(got_a, want_a, atol=1e-4) and torch.allclose(got_v, want_v, atol=1e-4)


def test_rollout_and_update_run_and_learn_signal_is_finite():
    import twoarmy_b200 as pkg
    P = _ppo()
    torch.manual_seed(0)
    agent = P.PPO(device="cuda:0")
    env = pkg.TwoarmyVecEnv(4, 256, 17, seed=1, autoreset=False)
    roll = P.VecRollout(env, agent, 64)
    buf = roll.collect()
    before = [p.detach().clone() for p in agent.actor.parameters()]
    al, vl = agent.update(P.with_her(buf, seed=3), minibatch=2048, epochs=1)
    assert np.isfinite(al) and np.isfinite(vl)
    assert any(not torch.equal(a, b) for a, b in zip(before, agent.actor.parameters()))
    assert buf.ended[:64].sum() > 0 and float(buf.r[:64].mean()) < 0


def _flat_params(net):
    return torch.cat([p.detach().float().reshape(-1) for p in net.parameters()])


def _fixture_update(P, fx, device, autocast, use_graph, two_streams=True):
    """PPO.update on the reference's own buffer (tests/golden/ppo_ref.npz) with the reference's sampler stream
    (torch.manual_seed(1) + SubsetRandomSampler on the CPU generator, exactly tests/test_ppo_cpu.py)."""
    torch.manual_seed(0)
    agent = P.PPO(device=device, autocast=autocast)
    agent.K_epochs, agent.batch_size = 2, 32
    agent.use_graph, agent.two_streams = use_graph, two_streams
    before = (_flat_params(agent.actor).cpu(), _flat_params(agent.critic).cpu())
    buf = {k: torch.from_numpy(fx[f"buf_{k}"]) for k in ("s", "a", "p", "g", "r", "a_logp")}
    torch.manual_seed(1)
    agent.update(buf, sampler_generator=torch.default_generator)
    return agent, before


def test_gpu_update_matches_reference_fixture_fp32(golden):
    """End-to-end update parity on the GPU, fp32 (autocast off, TF32 off): the production plumbing -- device buffers, the
    two-stream actor / critic fork-join, flat gradient buffers, fused capturable Adam and the CUDA-graph replay of the
    step -- must reproduce the REFERENCE's post-update parameter sums (soa/agent/PPO.py:103-158 run by
    tests/golden/make_golden_ppo.py) to fp32 round-off.  Tolerance: 2e-4 absolute on per-tensor sums (fp32 reductions in a
    different order over up to 590k elements; the CPU mirror holds 1e-5)."""
    P = _ppo()
    fx = golden("ppo_ref.npz")
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        agent, _ = _fixture_update(P, fx, "cuda:0", autocast=False, use_graph=True)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf
    got_a = np.array([p.detach().double().sum().item() for p in agent.actor.parameters()])
    got_c = np.array([p.detach().double().sum().item() for p in agent.critic.parameters()])
    np.testing.assert_allclose(got_a, fx["upd_actor_sums"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(got_c, fx["upd_critic_sums"], rtol=2e-4, atol=2e-4)
    assert agent.update_count == 2 * 3
    for name, net in (("actor", agent.actor), ("critic", agent.critic)):   # gradients still live inside the flat buffer
        flat = agent._flat[name]
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * flat.element_size()
        assert all(lo <= p.grad.data_ptr() < hi for p in net.parameters())


def test_gpu_update_bf16_fused_path_tracks_reference_fixture(golden):
    """The PRODUCTION update (bf16 autocast, tcgen05 first layer, parity-plane data gradients, CUDA-graph replay) on the
    reference's buffer against the fp32 CPU mirror that tests/test_ppo_cpu.py pins to the reference.  bf16 activations move
    individual Adam steps (sign-like at the first steps), so the comparison is on the parameter DELTAS of the 6 optimiser
    steps: per network, cosine similarity >= 0.90 and |sum(delta_gpu) - sum(delta_ref)| <= 5 % of sum |delta_ref|."""
    P = _ppo()
    fx = golden("ppo_ref.npz")
    gpu, gb = _fixture_update(P, fx, "cuda:0", autocast=True, use_graph=True)
    cpu, cb = _fixture_update(P, fx, "cpu", autocast=False, use_graph=False)
    for name, i in (("actor", 0), ("critic", 1)):
        net_g, net_c = getattr(gpu, name), getattr(cpu, name)
        dg = _flat_params(net_g).cpu() - gb[i]
        dc = _flat_params(net_c) - cb[i]
        assert torch.equal(gb[i], cb[i])                      # same initial weights
        cos = float(torch.dot(dg, dc) / (dg.norm() * dc.norm()))
        rel = float((dg.sum() - dc.sum()).abs() / dc.abs().sum())
        print(f"{name}: cosine {cos:.4f}, relative sum error {rel:.5f}, max |delta| {float(dg.abs().max()):.2e}")
        assert cos >= 0.90, (name, cos)
        assert rel <= 0.05, (name, rel)
        assert float(dg.abs().max()) <= 6 * 1.05e-4           # 6 Adam steps of lr 1e-4


@pytest.mark.parametrize("autocast", [True])
def test_graph_replay_equals_eager_steps(golden, autocast):
    """The CUDA-graph replay of the PRODUCTION optimiser step (bf16 autocast, fused kernels) vs the same steps launched
    eagerly (same weights, same minibatches).  (The fp32 debugging path goes through cuDNN's fp32 convolutions, whose
    algorithm choice differs between capture and eager launches; it is pinned to the reference fixture instead,
    test_gpu_update_matches_reference_fixture_fp32.)
    Every kernel of the step is deterministic except the float atomics that finish the conv1 weight gradient and the
    channel sums (cross-CTA accumulation order), so gradients agree to ~1e-6 relative and the first Adam steps, which
    normalise each gradient element by its own magnitude, can amplify that only where |g| is at round-off level:
    >= 99.9 % of all parameters bit-identical, none further apart than two Adam steps (2.1e-4)."""
    P = _ppo()
    fx = golden("ppo_ref.npz")
    outs = []
    for use_graph in (True, False):
        agent, _ = _fixture_update(P, fx, "cuda:0", autocast=autocast, use_graph=use_graph)
        outs.append((torch.cat([_flat_params(agent.actor), _flat_params(agent.critic)]), agent.last_action_loss, agent.last_value_loss))
    a, b = outs[0][0], outs[1][0]
    same = float((a == b).float().mean())
    print(f"autocast={autocast}: bit-identical {same:.6f}, max diff {float((a - b).abs().max()):.3e}")
    assert same >= 0.999 and float((a - b).abs().max()) <= 2.1e-4
    assert outs[0][1] == pytest.approx(outs[1][1], rel=1e-3, abs=1e-5) and outs[0][2] == pytest.approx(outs[1][2], rel=1e-3, abs=1e-5)


def _random_buffer(n, seed):
    g = torch.Generator().manual_seed(seed)
    codes = torch.tensor([0, 1, 2, 4], dtype=torch.uint8)[torch.randint(0, 4, (n, 5, 289), generator=g)]
    return {"s": codes, "p": torch.randint(1, 16, (n, 5, 2), generator=g).float(), "a": torch.randint(0, 5, (n, 1), generator=g),
            "g": torch.tensor([[2.0, 14.0]]).repeat(n, 1), "r": (torch.rand(n, 1, generator=g) - 0.5) * 0.2,
            "a_logp": torch.log(torch.rand(n, 1, generator=g) * 0.3 + 0.1)}


@pytest.mark.parametrize("n", [512, 300])
def test_fused_step_gradients_match_autograd_step(n):
    """fused_step.FusedNet (explicit forward / backward, analytic loss gradients, deterministic bias sums) on one minibatch
    against (a) the fp32 autograd step (TF32 off) and (b) the bf16 autograd step it replaces, same weights.  bf16
    activations put several per cent of noise on the early layers' gradients of a 512-sample minibatch (measured: 9-11 %
    on conv1, 1-3 % on the head for BOTH bf16 paths), so the bar is relative: for every parameter the fused step's
    distance to the fp32 gradient is at most 1.5 x the bf16 autograd step's (+ 1 % of the gradient norm), and the losses
    agree to 1e-3.  lr = 0 keeps the gradients in the flat buffers."""
    P = _ppo()
    buf = {k: v.cuda() for k, v in _random_buffer(n, 7).items()}
    grads, losses = {}, {}
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for mode in ("fp32", "autograd", "fused"):
            torch.manual_seed(0)
            agent = P.PPO(device="cuda:0", autocast=mode != "fp32")
            agent.fused_step = mode == "fused"
            # (no clipping here: a sample whose ratio sits at the clip boundary contributes all or nothing, so a bf16-sized
            # change of one logit moves the whole gradient by 1/B -- the clipped branch is covered by the update tests)
            agent.clip_param = 1e9
            for opt in (agent.optimizer_actor, agent.optimizer_critic):
                opt.param_groups[0]["lr"] = 0.0
            step, B, bs, _ = agent._make_step(buf, minibatch=n)
            la, lc = step(torch.arange(n, device="cuda"))
            torch.cuda.synchronize()
            assert (agent._fused is not None) == (mode == "fused")
            losses[mode] = (float(la), float(lc))
            grads[mode] = [p.grad.detach().float().clone() for net in (agent.actor, agent.critic) for p in net.parameters()]
            names = [f"{nn}.{k}" for nn, net in (("actor", agent.actor), ("critic", agent.critic)) for k, _ in net.named_parameters()]
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf
    print("losses", losses)
    for k in (0, 1):   # (unclipped ratios reach 10: the bf16 forward moves the mean loss by up to ~1 %)
        assert losses["fused"][k] == pytest.approx(losses["autograd"][k], rel=1e-2, abs=1e-3), losses
        assert losses["fused"][k] == pytest.approx(losses["fp32"][k], rel=5e-2, abs=5e-3), losses
    bad = {}
    for name, ref, a, f in zip(names, grads["fp32"], grads["autograd"], grads["fused"]):
        nr = max(float(ref.norm()), 1e-12)
        ea, ef = float((a - ref).norm()) / nr, float((f - ref).norm()) / nr
        print(f"{name:40s} |g| {nr:.3e}   bf16 autograd vs fp32 {ea:.4f}   fused vs fp32 {ef:.4f}")
        if ef > 1.5 * ea + 1e-2:
            bad[name] = (round(ea, 4), round(ef, 4))
    assert not bad, bad


def test_fused_step_is_bitwise_reproducible(golden):
    """Every reduction of the hand-scheduled step except the conv1 weight gradient's final float atomics has a fixed order:
    two updates from the same state give the same parameters except where |gradient| is at round-off level
    (>= 99.99 % of the elements bit-identical), and checkpoints carry the Adam moments in torch.optim.Adam's format."""
    P = _ppo()
    fx = golden("ppo_ref.npz")
    outs = []
    for _ in range(2):
        agent, _ = _fixture_update(P, fx, "cuda:0", autocast=True, use_graph=True)
        outs.append(torch.cat([_flat_params(agent.actor), _flat_params(agent.critic)]))
    assert float((outs[0] == outs[1]).float().mean()) >= 0.9999
    sd = agent.state_dict()
    st = sd["optimizer_actor"]["state"]
    assert len(st) == 16 and float(st[0]["step"]) == 6.0 and st[2]["exp_avg"].shape == (64, 64, 3, 3)
    other = P.PPO(device="cuda:0")
    other.load_state_dict(sd)
    assert torch.equal(_flat_params(other.actor), _flat_params(agent.actor))


def test_conv1_weight_gradient_run_to_run_bound():
    """conv1_bwd_tc_kernel finishes with one float atomicAdd per value and CTA: the accumulation order over CTAs is not
    fixed, so the gradient is reproducible only to fp32 round-off.  Documented bound: 1e-5 of the largest entry."""
    import twoarmy_b200 as pkg
    P = _ppo()
    C1 = importlib.import_module(pkg.__name__ + ".conv1")
    torch.manual_seed(0)
    conv = P.TINet().cuda().cnn_base[0]
    g = torch.Generator().manual_seed(2)
    codes = torch.tensor([0, 1, 2, 4], dtype=torch.uint8)[torch.randint(0, 4, (2048, 5, 289), generator=g)].cuda()
    gy = torch.randn((2048, 64, 33, 33), generator=torch.Generator().manual_seed(3)).cuda().to(torch.bfloat16)
    grads = []
    for _ in range(3):
        conv.weight.grad = None; conv.bias.grad = None
        y = C1.conv1_relu(codes[:, 1:5], conv)
        y.backward(gy)
        grads.append((conv.weight.grad.clone(), conv.bias.grad.clone()))
    for gw, gb in grads[1:]:
        assert float((gw - grads[0][0]).abs().max()) <= 1e-5 * float(grads[0][0].abs().max())
        assert float((gb - grads[0][1]).abs().max()) <= 1e-5 * float(grads[0][1].abs().max())


@pytest.fixture(params=["tcgen05-ws", "tcgen05", "fma"])
def conv1_kernel(request):
    """Run the test once through each kernel set of the fused first layer: the warp-specialised tcgen05 forward with
    tensor-map stores (default) / the single-role tcgen05 forward (TA_CONV1_TC=1) / the FP32-FMA forward, each with the matching
    weight-gradient kernel (tcgen05 for the first two)."""
    import twoarmy_b200 as pkg
    L = pkg._capi.lib()
    on = {"tcgen05-ws": 2, "tcgen05": 1, "fma": 0}[request.param]
    prev, prev_bwd = L.ta_debug_conv1_tc(on), L.ta_debug_conv1_bwd_tc(1 if on else 0)
    yield request.param
    assert L.ta_debug_conv1_tc_failed() == 0          # no tcgen05 launch gave up on its MMA barrier
    L.ta_debug_conv1_tc(prev)
    L.ta_debug_conv1_bwd_tc(1 if prev_bwd != 0 else 0)


@pytest.mark.parametrize("dtype", ["u8", "f32"])
def test_fused_conv1_matches_cudnn_layer(dtype, conv1_kernel):
    """ta_conv1_fwd / ta_conv1_bwd (tcgen05 kernels with bf16 hi/lo split inputs, and the FP32-FMA kernels) ==
    decode + UpsamplingNearest2d(4) + Conv2d(4,64,4,2) + ReLU evaluated in FLOAT64 on the CPU (no TF32, no library
    algorithm choice).  Forward: the bf16 output is the correctly rounded value up to the kernel's fp32-grade
    accumulation error -- half a bf16 ulp (2^-8 relative) + 1e-4 absolute (hi*hi + hi*lo + lo*hi drops the lo*lo
    terms: 2^-16 of the summed magnitudes).  Weight / bias gradients: 1e-3 of the largest entry."""
    import twoarmy_b200 as pkg
    P = _ppo()
    C1 = importlib.import_module(pkg.__name__ + ".conv1")
    torch.manual_seed(0)
    net = P.TINet().cuda()
    conv = net.cnn_base[0]
    g = torch.Generator().manual_seed(2)
    B = 37 if dtype == "u8" else 700        # 700 x 289 positions: more 128-row tiles than resident CTAs
    codes = torch.tensor([0, 1, 2, 4], dtype=torch.uint8)[torch.randint(0, 4, (B, 5, 289), generator=g)].cuda()
    x_in = codes[:, 1:5] if dtype == "u8" else P.decode_matrix(codes[:, 1:5]).contiguous()
    # reference in float64 on the CPU: no TF32, no cuDNN algorithm choice involved
    conv64 = torch.nn.Conv2d(4, 64, 4, 2).double()
    conv64.weight.data.copy_(conv.weight.detach().cpu().double()); conv64.bias.data.copy_(conv.bias.detach().cpu().double())
    xf = P.decode_matrix(codes[:, 1:5]).view(B, 4, 17, 17).cpu().double()
    want = torch.relu(conv64(torch.nn.functional.interpolate(xf, scale_factor=4, mode="nearest")))
    gy = torch.randn(want.shape, generator=torch.Generator().manual_seed(3))
    got = C1.conv1_relu(x_in, conv)
    assert got.shape == want.shape and got.dtype == torch.bfloat16
    # forward: one bf16 rounding of the fp32-grade result -> within 1 bf16 ulp (2^-8 relative) of the float64 value
    err = (got.double().cpu() - want).abs()
    assert bool((err <= want.abs() * (1.02 * 2.0 ** -8) + 1e-4).all()), float((err - want.abs() * 2.0 ** -8).max())
    # gradients for the same upstream gradient AND the same ReLU mask (the kernel's own output > 0) on both sides: the
    # reference differentiates the linear layer only, so that pre-activations within round-off of zero -- where a float64
    # and an fp32-grade evaluation may disagree about the sign -- do not enter the comparison
    mask = (got.detach().cpu() > 0).double()
    conv.weight.grad = None; conv.bias.grad = None
    got.backward(gy.cuda().to(torch.bfloat16))
    gy16 = gy.to(torch.bfloat16).double()           # the kernel consumes the bf16-rounded upstream gradient
    (conv64(torch.nn.functional.interpolate(xf, scale_factor=4, mode="nearest")) * (gy16 * mask)).sum().backward()
    gw, gb = conv64.weight.grad, conv64.bias.grad
    ew = float((conv.weight.grad.double().cpu() - gw).abs().max()) / float(gw.abs().max())
    eb = float((conv.bias.grad.double().cpu() - gb).abs().max()) / float(gb.abs().max())
    print(f"conv1 {conv1_kernel} {dtype}: weight-gradient error {ew:.2e}, bias-gradient error {eb:.2e} (relative to the largest entry)")
    assert ew <= 1e-3 and eb <= 1e-3, (ew, eb)


@pytest.mark.parametrize("layer", [2, 4])
def test_conv_s2_gemm_dgrad_matches_cudnn(layer):
    """conv_s2_relu (cuDNN fused forward, cuDNN wgrad, channel-sum bias gradient, GEMM + ta_col2im_s2 data
    gradient) == relu(nn.Conv2d) autograd in bf16."""
    import twoarmy_b200 as pkg
    P = _ppo()
    C1 = importlib.import_module(pkg.__name__ + ".conv1")
    torch.manual_seed(0)
    conv = P.TINet().cuda().cnn_base[layer]
    cin, hin = (64, 33) if layer == 2 else (64, 16)
    g = torch.Generator().manual_seed(5)
    x0 = torch.randn(19, cin, hin, hin, generator=g).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    outs = []
    for mode in ("ref", "gemm"):
        x = x0.clone().requires_grad_(True)
        conv.weight.grad = None; conv.bias.grad = None
        if mode == "ref":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = torch.relu(conv(x))
        else:
            y = C1.conv_s2_relu(x, conv)
        gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(6)).cuda().to(y.dtype)
        (y.float() * gy.float()).sum().backward()
        outs.append((y.float(), x.grad.float(), conv.weight.grad.clone(), conv.bias.grad.clone()))
    # the fused forward rounds conv + bias once, the unfused one twice: activations within one bf16 ulp of
    # zero can land on either side of the ReLU, so single gradient entries may differ by one term; compare
    # y elementwise to rounding and the gradients in the Frobenius norm
    y0, y1 = outs[0][0], outs[1][0]
    assert float((y0 - y1).abs().max()) <= 2e-2 * float(y0.abs().max())
    for name, a, b in zip(("dx", "dw", "db"), outs[0][1:], outs[1][1:]):
        assert float((a - b).norm()) <= 2e-2 * float(a.norm()), name


@pytest.mark.parametrize("shape", [(64, 64, 3), (128, 64, 4), (8, 16, 3)])
def test_parity_class_weights_and_merged_planes(shape):
    """ta_parity_class_weights (one launch) == slicing / flipping / padding the conv weight in torch (contiguous and
    channels-last input), and the merged-plane convolution built from it == conv_transpose2d (the data gradient)."""
    import twoarmy_b200 as pkg
    import torch.nn.functional as F
    C1 = importlib.import_module(pkg.__name__ + ".conv1")
    cout, cin, k = shape
    w = (torch.randn((cout, cin, k, k), generator=torch.Generator().manual_seed(4)) * 0.1).cuda().to(torch.bfloat16)
    want = C1.parity_class_weights(w.float())                       # the torch path (not bf16 -> no kernel)
    for wv in (w, w.contiguous(memory_format=torch.channels_last)):
        got = C1.parity_class_weights(wv)
        assert got.shape == want.shape and got.is_contiguous(memory_format=torch.channels_last)
        assert torch.equal(got.float(), want)
    oh = 6
    dz = torch.randn((3, cout, oh, oh), generator=torch.Generator().manual_seed(5)).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    planes = C1._class_planes(dz, w).float()                        # [B, oh+1, oh+1, 4*cin]
    dx = F.conv_transpose2d(dz.float(), w.float(), stride=2)        # [B, cin, 2*oh+k-2, ...]
    H = dx.shape[2]
    for c in range(4):
        pa, pb = c >> 1, c & 1
        want_c = dx[:, :, pa::2, pb::2].permute(0, 2, 3, 1)
        got_c = planes[:, :want_c.shape[1], :want_c.shape[2], c * cin:(c + 1) * cin]
        assert float((got_c - want_c).abs().max()) <= 2e-2 * float(want_c.abs().max()) + 1e-3, (c, H)


@pytest.mark.parametrize("dtype,depth", [("u8", 2), ("f32", 2), ("u8", 3)])
def test_stem_parity_plane_dgrad_matches_unfused_layers(dtype, depth):
    """conv1._Stem (conv1 + conv2 [+ conv3] in one autograd node; data gradients as four parity-class stride-1
    convolutions, conv3's interleaved by ta_planes_to_dense_relu, conv2's read in place by ta_conv1_bwd_planes) ==
    conv1_relu followed by conv_s2_relu (GEMM + col2im data gradient, ta_conv1_bwd): same forward bit for bit,
    parameter gradients to bf16 rounding of the intermediates."""
    import twoarmy_b200 as pkg
    P = _ppo()
    C1 = importlib.import_module(pkg.__name__ + ".conv1")
    torch.manual_seed(0)
    net = P.TINet().cuda()
    conv1, conv2, conv3 = net.cnn_base[0], net.cnn_base[2], net.cnn_base[4]
    g = torch.Generator().manual_seed(2)
    B = 301                                   # 301 * 289 positions: ragged last tile, several tiles per CTA
    codes = torch.tensor([0, 1, 2, 4], dtype=torch.uint8)[torch.randint(0, 4, (B, 5, 289), generator=g)].cuda()
    x_in = codes[:, 1:5] if dtype == "u8" else P.decode_matrix(codes[:, 1:5]).contiguous()
    gy = None
    outs = []
    for mode in ("layers", "stem"):
        prms = (conv1.weight, conv1.bias, conv2.weight, conv2.bias) + ((conv3.weight, conv3.bias) if depth == 3 else ())
        for prm in prms:
            prm.grad = None
        if mode == "stem":
            y2 = C1.stem_relu(x_in, conv1, conv2, conv3 if depth == 3 else None)
        else:
            y2 = C1.conv_s2_relu(C1.conv1_relu(x_in, conv1), conv2)
            y2 = C1.conv_s2_relu(y2, conv3) if depth == 3 else y2
        if gy is None:
            gy = torch.randn(y2.shape, generator=torch.Generator().manual_seed(3)).cuda().to(torch.bfloat16)
        (y2.float() * gy.float()).sum().backward()
        outs.append((y2.detach().float().clone(), [prm.grad.detach().clone() for prm in prms]))
    assert torch.equal(outs[0][0], outs[1][0])
    for name, a, b in zip(("conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias"), outs[0][1], outs[1][1]):
        assert float((a - b).norm()) <= 1e-2 * float(a.norm()), name
    assert pkg._capi.lib().ta_debug_conv1_tc_failed() == 0


@pytest.mark.parametrize("shape", [(5, 7, 7, 128, 3), (3, 33, 33, 64, 3), (2, 16, 16, 64, 4), (1, 9, 8, 8, 3)])
def test_im2col_s2_kernel_is_the_unfold_and_col2im_its_adjoint(shape):
    """ta_im2col_s2 == F.unfold(kernel k, stride 2) with columns ordered (ky, kx, c), bit for bit (it only moves
    bf16 values); ta_col2im_s2 is its adjoint: <im2col(x), d> == <x, col2im(d)> in fp32 up to bf16 rounding."""
    import twoarmy_b200 as pkg
    import torch.nn.functional as F
    C1 = importlib.import_module(pkg.__name__ + ".conv1")
    B, H, W, Cc, k = shape
    g = torch.Generator().manual_seed(11)
    x = torch.randn((B, H, W, Cc), generator=g).cuda().to(torch.bfloat16).requires_grad_(True)
    cols = C1.im2col_s2(x, None, k)                                       # [B, OH*OW*k*k, C]
    OH, OW = (H - k) // 2 + 1, (W - k) // 2 + 1
    want = F.unfold(x.detach().float().permute(0, 3, 1, 2), k, stride=2)  # [B, C*k*k, OH*OW], rows (c, ky, kx)
    want = want.view(B, Cc, k * k, OH * OW).permute(0, 3, 2, 1).reshape(B, OH * OW * k * k, Cc)
    assert torch.equal(cols.detach().float(), want)
    d = torch.randn(cols.shape, generator=g).cuda().to(torch.bfloat16)
    (cols.float() * d.float()).sum().backward()
    want_dx = F.fold(d.float().view(B, OH * OW, k * k, Cc).permute(0, 3, 2, 1).reshape(B, Cc * k * k, OH * OW), (H, W), k, stride=2)
    got_dx = x.grad.float().permute(0, 3, 1, 2)
    assert float((got_dx - want_dx).abs().max()) <= 2e-2 * float(want_dx.abs().max())


@pytest.mark.parametrize("dims", [(10, 128), (2304, 256), (384, 512)])
def test_linear_relu_matches_autograd(dims):
    """conv1.linear_relu (bias + ReLU in the GEMM epilogue, channel-sum bias gradient) == relu(F.linear) autograd in bf16."""
    import twoarmy_b200 as pkg
    C1 = importlib.import_module(pkg.__name__ + ".conv1")
    fin, fout = dims
    torch.manual_seed(1)
    lin = torch.nn.Linear(fin, fout).cuda()
    x0 = torch.randn(517, fin, device="cuda").to(torch.bfloat16)
    gy = torch.randn(517, fout, device="cuda").to(torch.bfloat16)
    outs = []
    for mode in ("ref", "fused"):
        lin.weight.grad = None; lin.bias.grad = None
        x = x0.clone().requires_grad_(True)
        if mode == "ref":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = torch.relu(lin(x))
        else:
            y = C1.linear_relu(x, lin)
        (y.float() * gy.float()).sum().backward()
        outs.append((y.detach().float(), x.grad.float(), lin.weight.grad.clone(), lin.bias.grad.clone()))
    assert float((outs[0][0] - outs[1][0]).abs().max()) <= 2e-2 * float(outs[0][0].abs().max())
    for name, a, b in zip(("dx", "dw", "db"), outs[0][1:], outs[1][1:]):
        assert float((a - b).norm()) <= 2e-2 * float(a.norm()), name


def test_predictor_agent_rollout_and_update_on_gpu():
    """BASELINE configs[4] plumbing: ppo_predictor drives VecRollout and updates on the device."""
    import twoarmy_b200 as pkg
    P = _ppo()
    M = importlib.import_module(pkg.__name__ + ".predictor")
    torch.manual_seed(0)
    agent = M.ppo_predictor(device="cuda:0")
    env = pkg.TwoarmyVecEnv(4, 128, 17, seed=2, autoreset=False)
    roll = P.VecRollout(env, agent, 16)
    buf = roll.collect()
    before = [p.detach().clone() for p in agent.critic.parameters()]
    al, vl = agent.update(buf.flat(), minibatch=1024, epochs=1)
    assert np.isfinite(al) and np.isfinite(vl)
    assert any(not torch.equal(a, b) for a, b in zip(before, agent.critic.parameters()))
    # 8 full minibatches: the third one on is replayed from the CUDA graph of the step (predictor forward included)
    al, vl = agent.update(buf.flat(), minibatch=256, epochs=1)
    assert np.isfinite(al) and np.isfinite(vl)
    # GPU (bf16) prediction stays close to the fp32 CPU one for the same weights
    x = P.decode_matrix(buf.s[3, :8, 0:4]).float()
    got = agent.pred_states(x)[0].cpu()
    cpu = M.ppo_predictor(device="cpu", autocast=False)
    cpu.load_predictor(agent.state_dict())
    want = cpu.pred_states(x.cpu())[0]
    assert float((got - want).abs().max()) < 0.05 * max(1.0, float(want.abs().max()))
    idx, e = M.pre_transition_records(buf.ended[:16])
    assert idx.shape[1] == 9 and idx.shape[0] == e.shape[0]


def test_step_graph_kept_across_updates_equals_recapture():
    """The optimiser-step graph captured by the first update() is replayed by later ones on the same rollout buffer
    (PPO.keep_graph); the small per-update tensors it reads (advantages, targets, positions ...) are persistent copies that
    every update() refreshes.  Three updates, rewards and frames rewritten in place in between, against re-capturing in
    every update.  At this size two runs of the SAME mode already differ in most parameter bits (conv1's float atomics and
    cuDNN's split-K weight gradients accumulate in a run-dependent order, Adam amplifies round-off on near-zero gradients;
    measured <= 1e-3 absolute after three updates, scripts/probe_keep_graph.py), so the comparison is by the losses of every
    update (a stale advantage / target tensor would reproduce the previous update's losses instead) and a parameter bound."""
    P = _ppo()
    dev = torch.device("cuda:0")
    B, mb = 2048, 512
    outs, losses, reused = [], [], []
    for keep in (True, False):
        torch.manual_seed(0)
        agent = P.PPO(device=dev)
        agent.K_epochs, agent.keep_graph = 2, keep
        g = torch.Generator(device=dev).manual_seed(1)
        buf = {"s": torch.randint(0, 3, (B, 5, 289), generator=g, device=dev, dtype=torch.uint8),
               "p": torch.randint(1, 16, (B, 5, 2), generator=g, device=dev).float(),
               "a": torch.randint(0, 5, (B, 1), generator=g, device=dev), "g": torch.tensor([[2.0, 14.0]], device=dev).repeat(B, 1),
               "r": torch.rand(B, 1, generator=g, device=dev) - 0.5, "a_logp": torch.log(torch.rand(B, 1, generator=g, device=dev) * 0.3 + 0.1)}
        n, ls = 0, []
        for it in range(3):
            torch.manual_seed(10 + it)
            ls.append(agent.update(buf, minibatch=mb))
            n += int(agent._graph_cache is not None)
            buf["r"].copy_((torch.rand(B, 1, generator=g, device=dev) - 0.5) * (it + 2))      # same storage, new values
            buf["s"][:, :, ::7] = (buf["s"][:, :, ::7] + 1) % 3
        reused.append(n)
        losses.append(np.array(ls))
        outs.append(torch.cat([_flat_params(agent.actor), _flat_params(agent.critic)]))
    assert reused == [3, 0]
    np.testing.assert_allclose(losses[0], losses[1], rtol=3e-2, atol=2e-3)
    # the updates are not interchangeable: the value loss follows the rescaled rewards (what a stale target would miss)
    assert losses[1][2, 1] > 1.5 * losses[1][0, 1]
    assert float((outs[0] - outs[1]).abs().max()) < 5e-3


def test_next_value_shared_with_next_record():
    """RolloutBuffer.flat()'s `next_same` hint: where the episode did not end at step t, record t + 1 of the same env holds
    record t's frames 1..4 / positions 1..4 as ITS frames 0..3 / positions 0..3 (checked here on a real rollout with episode
    ends), so PPO.values reads V(s') from the first critic pass.  The values are those of the two full passes of the
    reference (PPO.py:113-114): identical where shared (same kernels on identical inputs in equally sized chunks), within
    bf16 rounding for the remaining samples (evaluated in one smaller batch), and an update() gives the same losses."""
    import twoarmy_b200 as pkg
    P = _ppo()
    torch.manual_seed(0)
    agent = P.PPO(device="cuda:0")
    env = pkg.TwoarmyVecEnv(4, 512, 17, seed=5, autoreset=False)
    T, N = 96, 512
    roll = P.VecRollout(env, agent, T)
    buf = roll.collect()
    flat = buf.flat()
    nxt = flat["next_same"].view(T, N)
    assert int(buf.ended[:T].sum()) > 0 and not bool(nxt[T - 1].any())
    assert bool((nxt[:T - 1] == (buf.ended[:T - 1] == 0)).all())
    same = nxt[:T - 1]
    assert bool((buf.s[1:T][same][:, 0:4] == buf.s[:T - 1][same][:, 1:5]).all())
    assert bool((buf.p[1:T][same][:, 0:4] == buf.p[:T - 1][same][:, 1:5]).all())
    # and where the episode ended the next record is NOT the continuation (so the hint is needed, not just harmless)
    ended = ~same
    assert bool((buf.s[1:T][ended][:, 0:4] != buf.s[:T - 1][ended][:, 1:5]).flatten(1).any(1).any())
    with torch.no_grad():
        v2, vn2 = agent.values(flat["s"], flat["p"], flat["g"])
        v1, vn1 = agent.values(flat["s"], flat["p"], flat["g"], next_same=flat["next_same"], next_stride=flat["next_stride"])
    assert torch.equal(v1, v2)
    shared = flat["next_same"]
    assert float((vn1[shared] == vn2[shared]).float().mean()) >= 0.99
    scale = float(vn2.abs().max())
    assert float((vn1 - vn2).abs().max()) <= 2e-2 * scale + 1e-3
    losses = []
    for share in (True, False):
        torch.manual_seed(0)
        ag = P.PPO(device="cuda:0")
        ag.share_next_value, ag.K_epochs = share, 1
        torch.manual_seed(3)
        losses.append(ag.update(flat, minibatch=4096))
    np.testing.assert_allclose(losses[0], losses[1], rtol=2e-2, atol=1e-3)


def test_lstm_gates_kernel_matches_torch_cell():
    """ta_lstm_gates == torch.nn.LSTM's cell arithmetic (gate order i, f, g, o) on the same fp32 pre-activations:
    c to fp32 rounding, h to the rounding of its bf16 output; with and without the second gate operand, and with the h
    output written into the right half of a wider [x | h] operand (strided rows)."""
    import ctypes as C
    import twoarmy_b200 as pkg
    L = pkg._capi.lib()
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(5)
    for B, H, wide in ((3, 1024, False), (257, 1024, True), (64, 8, False)):
        gx = torch.randn((B, 4 * H), device=dev, generator=g) * 2
        gh = torch.randn((B, 4 * H), device=dev, generator=g)
        bias = torch.randn(4 * H, device=dev, generator=g) * 0.1
        c0 = torch.randn((B, H), device=dev, generator=g)
        for use_gh in (True, False):
            c = c0.clone()
            buf = torch.zeros((B, 2 * H if wide else H), dtype=torch.bfloat16, device=dev)
            h_out = buf[:, -H:]
            pkg._capi.check(L.ta_lstm_gates(C.c_void_p(gx.data_ptr()), C.c_void_p(gh.data_ptr()) if use_gh else None,
                                            C.c_void_p(bias.data_ptr()), C.c_void_p(c.data_ptr()), C.c_void_p(h_out.data_ptr()),
                                            buf.stride(0), B, H, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "ta_lstm_gates")
            pre = (gx + (gh if use_gh else 0) + bias).double()
            i, f, gg, o = pre.split(H, 1)
            c_want = torch.sigmoid(f) * c0.double() + torch.sigmoid(i) * torch.tanh(gg)
            h_want = torch.sigmoid(o) * torch.tanh(c_want)
            assert float((c.double() - c_want).abs().max()) < 2e-6 * max(1.0, float(c_want.abs().max()))
            assert float((h_out.double() - h_want).abs().max()) < 2 ** -8          # |h| < 1: one bf16 ulp
            if wide:
                assert float(buf[:, :H].abs().max()) == 0.0                         # the x half is not touched


def test_fast_lstm_forward_matches_nn_lstm():
    """predictor.LSTM's hand-scheduled forward (bf16 GEMMs with fp32 gates + ta_lstm_gates) against torch.nn.LSTM in fp32
    with the same weights: the 4 teacher-forced outputs and the 3 self-fed steps (all_net.py:76-98)."""
    import twoarmy_b200 as pkg
    M = importlib.import_module(pkg.__name__ + ".predictor")
    torch.manual_seed(3)
    m = M.LSTM().to("cuda:0")
    m.h_0, m.c_0 = m.h_0.to("cuda:0"), m.c_0.to("cuda:0")
    with torch.no_grad():
        for p in m.recurrent_model.parameters():     # (the default init is tiny; make the recurrence matter)
            p.mul_(3.0)
    z = torch.randn((37, 4, 64, 4, 4), device="cuda:0")
    with torch.no_grad():
        want, _ = m(z)                                                       # fp32, cuDNN (no autocast -> reference path)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            got, zc = m(z)                                                   # the fast path
    assert got.dtype == torch.bfloat16 and got.shape == want.shape == (37, 7, 64, 4, 4) and zc.shape == (37, 4, 1024)
    err = float((got.float() - want).abs().max())
    assert err < 0.03 * max(1.0, float(want.abs().max())), err
    # and its fp32 form (the same schedule with fp32 operands) agrees with cuDNN to fp32 rounding
    got32 = m._fast_forward(z.reshape(37, 4, 1024), torch.float32).reshape(37, 7, 64, 4, 4)
    assert float((got32 - want).abs().max()) < 1e-3 * max(1.0, float(want.abs().max()))      # (cuDNN's fp32 RNN may use TF32)


def test_prediction_table_equals_per_use_predictions():
    """ppo_predictor.update computes the frozen predictor's frames once per buffer row (_begin_update); what the networks
    receive from the table == what _cat computes from the frames themselves (bf16 GEMMs: equal up to the batch shape's
    rounding), and an update with the table tracks one without (TA_PRED_CACHE=0)."""
    import os
    import twoarmy_b200 as pkg
    P = _ppo()
    M = importlib.import_module(pkg.__name__ + ".predictor")
    torch.manual_seed(0)
    agent = M.ppo_predictor(device="cuda:0")
    env = pkg.TwoarmyVecEnv(4, 128, 17, seed=2, autoreset=False)
    buf = P.VecRollout(env, agent, 16).collect().flat()
    s = buf["s"].to("cuda:0")
    agent._begin_update(s)
    assert agent._pred_valid and agent._pred_table.shape == (s.shape[0], 4, 289)
    rows = torch.randperm(s.shape[0], device="cuda:0")[:300]
    got = agent._net_in(s[rows][:, 0:4], rows, 0)
    want = agent._cat(s[rows][:, 0:4])
    assert got.shape == want.shape == (300, 8, 289) and torch.equal(got[:, :4], want[:, :4])
    assert float((got - want).abs().max()) < 0.02 * max(1.0, float(want.abs().max()))
    agent._end_update()
    assert torch.equal(agent._net_in(s[rows][:, 0:4], rows, 0), want)     # table off: computed from the frames
    state = {k: v.clone() for k, v in agent.actor.state_dict().items()}
    outs = []
    for cache in ("1", "0"):
        os.environ["TA_PRED_CACHE"] = cache
        try:
            agent.actor.load_state_dict(state)
            torch.manual_seed(1)
            outs.append(agent.update(buf, minibatch=512, epochs=1))
        finally:
            os.environ.pop("TA_PRED_CACHE", None)
    assert all(np.isfinite(x) for o in outs for x in o)


def _predictor_agent_with_nontrivial_statistics():
    import twoarmy_b200 as pkg
    M = importlib.import_module(pkg.__name__ + ".predictor")
    torch.manual_seed(4)
    agent = M.ppo_predictor(device="cuda:0")
    g = torch.Generator(device="cuda:0").manual_seed(9)
    with torch.no_grad():
        for m in agent.encoder.cnn_base:
            if isinstance(m, torch.nn.BatchNorm2d):     # (fresh modules have mean 0 / var 1: make the fold matter)
                m.running_mean.copy_(torch.randn(m.num_features, device="cuda:0", generator=g) * 0.2)
                m.running_var.copy_(torch.rand(m.num_features, device="cuda:0", generator=g) + 0.5)
                m.weight.copy_(torch.rand(m.num_features, device="cuda:0", generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.num_features, device="cuda:0", generator=g) * 0.1)
        for net in (agent.encoder, agent.decoder):
            for m in net.cnn_base:
                if hasattr(m, "weight") and m.weight.dim() == 4:
                    m.weight.mul_(4.0)
                    m.bias.copy_(torch.randn(m.bias.shape, device="cuda:0", generator=g) * 0.1)
        for p in agent.predictor.recurrent_model.parameters():
            p.mul_(3.0)
    return agent, M


def test_fused_predictor_stacks_match_modules():
    """ta_pred_encoder / ta_pred_decoder (one fused kernel per stack, eval-mode BatchNorm and the average pool folded)
    against the torch modules in fp32: the encoder to the rounding of its bf16 output, the decoder -- fed the same bf16
    codes -- to fp32 accumulation order; uint8 codes and float LUT values give the same result."""
    import ctypes as C
    import twoarmy_b200 as pkg
    P = _ppo()
    agent, M = _predictor_agent_with_nontrivial_statistics()
    L = pkg._capi.lib()
    dev = "cuda:0"
    Mimg = 4 * 77
    codes = torch.tensor([0, 1, 2, 4], dtype=torch.uint8, device=dev)[torch.randint(0, 4, (Mimg, 289), device=dev)]
    e, d = agent._stack_arrays()
    ptr = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    agent.encoder.eval(); agent.decoder.eval()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            want_z = agent.encoder(P.decode_matrix(codes).float().view(Mimg, 1, 289))[0].reshape(Mimg, 1024)
            zs = []
            for x in (codes, P.decode_matrix(codes).float().contiguous()):
                z = torch.empty((Mimg, 1024), dtype=torch.bfloat16, device=dev)
                pkg._capi.check(L.ta_pred_encoder(ptr(x), 1 if x.dtype == torch.uint8 else 0, Mimg, ptr(e["w1"]), ptr(e["s1"]), ptr(e["t1"]),
                                                  ptr(e["w2"]), ptr(e["s2"]), ptr(e["t2"]), ptr(e["w3"]), ptr(e["s3"]), ptr(e["t3"]), ptr(z), st),
                                "ta_pred_encoder")
                zs.append(z)
            assert torch.equal(zs[0], zs[1])
            assert float(want_z.abs().max()) > 0.5
            err = (zs[0].float() - want_z).abs()
            assert float((err - 2 ** -8 * want_z.abs()).max()) < 1e-4, float(err.max())      # one bf16 ulp + fp32 noise
            zin = zs[0]
            want = agent.decoder(zin.float().view(Mimg, 1, 64, 4, 4))[0].reshape(Mimg, 289)
            out = torch.empty((Mimg, 289), dtype=torch.float32, device=dev)
            pkg._capi.check(L.ta_pred_decoder(ptr(zin), Mimg, ptr(d["w1"]), ptr(d["b1"]), ptr(d["w2"]), ptr(d["b2"]), ptr(d["w3"]),
                                              d["b3"], ptr(out), st), "ta_pred_decoder")
            assert float(want.abs().max()) > 0.1
            assert float((out - want).abs().max()) < 1e-4 * max(1.0, float(want.abs().max()))
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_fused_pred_states_tracks_fp32_modules():
    """pred_states on the GPU (fused stacks + hand-scheduled bf16 LSTM) against the unfused fp32 modules with the same
    weights (autocast off): bf16 rounding of the latent and of the LSTM's operands only."""
    agent, M = _predictor_agent_with_nontrivial_statistics()
    codes = torch.tensor([0, 1, 2, 4], dtype=torch.uint8, device="cuda:0")[torch.randint(0, 4, (50, 4, 289), device="cuda:0")]
    got = agent.pred_states(codes)[0]
    agent.autocast = False
    want = agent.pred_states(codes)[0]
    agent.autocast = True
    assert got.shape == want.shape == (50, 4, 289) and got.dtype == torch.float32
    assert float((got - want).abs().max()) < 0.03 * max(1.0, float(want.abs().max()))


def test_stem8_matches_unfused_layers():
    """The predictor agent's 8-channel TINet stem (conv1._Stem8: the folded first layer as two 4-channel passes joined by
    an addend, conv2 / conv3 with parity-plane data gradients, the first layer's weight gradient once per half) against the
    SAME network in fp32 (upsample + cuDNN layers, no autocast): its forward / gradient errors must not exceed those of
    the unfused bf16-autocast path (upsample + cuDNN in bf16, which also rounds the layer's input) by more than half
    plus 1 % -- two bf16 evaluations of one network differ from each other by several per cent (ReLUs near zero flip)."""
    import os
    import twoarmy_b200 as pkg
    P = _ppo()
    torch.manual_seed(0)
    net = P.TINet(in_frames=8).cuda()
    g = torch.Generator().manual_seed(2)
    B = 301
    codes = torch.tensor([0, 1, 2, 4], dtype=torch.uint8)[torch.randint(0, 4, (B, 4, 289), generator=g)].cuda()
    pred = (torch.rand((B, 4, 289), generator=g) * 1.8 - 0.9).cuda()
    x8 = torch.cat([P.decode_matrix(codes).float(), pred], 1).contiguous()
    pos = torch.rand((B, 4, 2), generator=g).cuda() * 16
    goal = torch.rand((B, 2), generator=g).cuda() * 16
    gy = None
    outs = {}
    prms = [net.cnn_base[i].weight for i in (0, 2, 4)] + [net.cnn_base[i].bias for i in (0, 2, 4)]
    names = ("conv1.weight", "conv2.weight", "conv3.weight", "conv1.bias", "conv2.bias", "conv3.bias")
    old_tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for mode in ("fp32", "bf16_unfused", "bf16_fused"):
            os.environ["TA_STEM8"] = "1" if mode == "bf16_fused" else "0"
            for prm in net.parameters():
                prm.grad = None
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode != "fp32"):
                y = net(x8, pos, goal)
            if gy is None:
                gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(3)).cuda()
            (y.float() * gy).sum().backward()
            outs[mode] = (y.detach().float().clone(), [prm.grad.detach().float().clone() for prm in prms])
    finally:
        os.environ.pop("TA_STEM8", None)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_tf32
    ref_y, ref_g = outs["fp32"]
    err = {m: {"forward": float((outs[m][0] - ref_y).norm() / ref_y.norm()),
               **{n: float((a - b).norm() / b.norm()) for n, a, b in zip(names, outs[m][1], ref_g)}} for m in ("bf16_unfused", "bf16_fused")}
    print("relative errors against fp32", err)
    for k, v in err["bf16_fused"].items():
        assert v <= 1.5 * err["bf16_unfused"][k] + 0.01, (k, err)
    assert pkg._capi.lib().ta_debug_conv1_tc_failed() == 0


def test_fused8_step_gradients_match_autograd_step():
    """The predictor agent under the hand-scheduled step (fused_step.FusedNet8: the 8-channel first layer as two passes of
    the folded 4-channel kernels, everything else as FusedNet) on one minibatch against the fp32 autograd step and the bf16
    autograd step (conv1._Stem8), same weights, same per-row predictions: the fused step's distance to the fp32 gradient
    is at most 1.5 x the bf16 autograd step's + 1 % for every parameter -- both halves of conv1's weight included."""
    import os
    import twoarmy_b200 as pkg
    M = importlib.import_module(pkg.__name__ + ".predictor")
    n = 512
    buf = {k: v.cuda() for k, v in _random_buffer(n, 7).items()}
    grads, losses = {}, {}
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for mode in ("fp32", "autograd", "fused"):
            torch.manual_seed(0)
            agent = M.ppo_predictor(device="cuda:0", autocast=mode != "fp32")
            with torch.no_grad():
                for prm in agent.predictor.recurrent_model.parameters():
                    prm.mul_(3.0)
                for net in (agent.encoder, agent.decoder):
                    for m in net.cnn_base:
                        if hasattr(m, "weight") and m.weight.dim() == 4:
                            m.weight.mul_(4.0)
            os.environ["TA_PPO_FUSED8"] = "1" if mode == "fused" else "0"
            agent.clip_param = 1e9
            for opt in (agent.optimizer_actor, agent.optimizer_critic):
                opt.param_groups[0]["lr"] = 0.0
            step, B, bs, _ = agent._make_step(buf, minibatch=n)
            if mode == "fp32":      # the other two read the bf16 table; give fp32 the same predictions (its own differ by bf16 rounding)
                pass
            la, lc = step(torch.arange(n, device="cuda"))
            torch.cuda.synchronize()
            assert (agent._fused is not None) == (mode == "fused")
            if mode == "fused":
                assert type(agent._fused["actor"]).__name__ == "FusedNet8"
            losses[mode] = (float(la), float(lc))
            grads[mode] = [p.grad.detach().float().clone() for net in (agent.actor, agent.critic) for p in net.parameters()]
            names = [f"{nn}.{k}" for nn, net in (("actor", agent.actor), ("critic", agent.critic)) for k, _ in net.named_parameters()]
            agent._end_update()
    finally:
        os.environ.pop("TA_PPO_FUSED8", None)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf
    print("losses", losses)
    for k in (0, 1):
        assert losses["fused"][k] == pytest.approx(losses["autograd"][k], rel=1e-2, abs=1e-3), losses
        assert losses["fused"][k] == pytest.approx(losses["fp32"][k], rel=5e-2, abs=5e-3), losses
    bad = {}
    for name, ref, a, f in zip(names, grads["fp32"], grads["autograd"], grads["fused"]):
        nr = max(float(ref.norm()), 1e-12)
        ea, ef = float((a - ref).norm()) / nr, float((f - ref).norm()) / nr
        print(f"{name:40s} |g| {nr:.3e}   bf16 autograd vs fp32 {ea:.4f}   fused vs fp32 {ef:.4f}")
        if ef > 1.5 * ea + 1e-2:
            bad[name] = (round(ea, 4), round(ef, 4))
    # both halves of the first layer's weight gradient individually (a missing or doubled half would hide in the norm above)
    for net_i in (0, 16):
        ref, f = grads["fp32"][net_i], grads["fused"][net_i]
        for half in (slice(0, 4), slice(4, 8)):
            assert float((f[:, half] - ref[:, half]).norm()) <= 0.25 * float(ref[:, half].norm()) + 1e-12, (net_i, half)
    assert not bad, bad
    assert pkg._capi.lib().ta_debug_conv1_tc_failed() == 0
