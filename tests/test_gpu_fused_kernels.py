"""GPU: the small kernels of the hand-scheduled optimiser step (csrc/ta_train.cuh) one by one against plain PyTorch
(fp32 / fp64 autograd of the same arithmetic, torch.optim.Adam), through the C ABI."""
import ctypes as C
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _lib():
    import twoarmy_b200 as pkg
    return pkg._capi.lib(), pkg._capi.check


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.mark.parametrize("rows,Cc,ld", [(4096, 512, 512), (4096, 256, 384), (4096, 128, 384), (36864, 256, 256), (200704, 128, 128),
                                         (1000, 64, 64), (37, 512, 512), (1, 64, 64)])
def test_relu_bwd_bias(rows, Cc, ld):
    """dz = dy * [y > 0] bit for bit, bias gradient = column sums of dz (fp32 accumulation, 1e-5 of the largest |sum| +
    the bf16 rounding of dz is already in dz), identical on every run (fixed reduction order), dy read through a row stride."""
    L, check = _lib()
    g = torch.Generator(device="cuda").manual_seed(rows + Cc)
    dyfull = torch.randn((rows, ld), generator=g, device="cuda").to(torch.bfloat16)
    off = ld - Cc
    dy = dyfull[:, off:]
    y = torch.relu(torch.randn((rows, Cc), generator=g, device="cuda")).to(torch.bfloat16)
    dz = torch.empty((rows, Cc), dtype=torch.bfloat16, device="cuda")
    n = int(L.ta_relu_bwd_bias_scratch_floats(rows, Cc))
    scratch = torch.zeros(n, dtype=torch.float32, device="cuda")
    outs = []
    for _ in range(3):
        db = torch.full((Cc,), float("nan"), device="cuda")
        check(L.ta_relu_bwd_bias(_p(dy), ld, _p(y), _p(dz), rows, Cc, _p(db), _p(scratch), _st()), "ta_relu_bwd_bias")
        outs.append(db.clone())
    want_dz = torch.where(y > 0, dy, torch.zeros_like(dy))
    assert torch.equal(dz, want_dz)
    want_db = want_dz.double().sum(0)
    assert float((outs[0].double() - want_db).abs().max()) <= 1e-5 * max(1.0, float(want_dz.double().abs().sum(0).max()))
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    # plain column sum (y == NULL)
    db = torch.empty((Cc,), device="cuda")
    check(L.ta_relu_bwd_bias(_p(dy), ld, None, None, rows, Cc, _p(db), _p(scratch), _st()), "ta_relu_bwd_bias")
    want = dy.double().sum(0)
    assert float((db.double() - want).abs().max()) <= 1e-5 * max(1.0, float(dy.double().abs().sum(0).max()))


@pytest.mark.parametrize("B,H,Cc", [(37, 16, 64), (5, 33, 64), (3, 7, 128)])
def test_planes_relu_bwd_bias(B, H, Cc):
    """ta_planes_relu_bwd_bias == ta_planes_to_dense_relu followed by the column sum: merged parity planes -> dense
    gradient masked with the ReLU of the layer below (bit for bit) + its bias gradient."""
    L, check = _lib()
    g = torch.Generator(device="cuda").manual_seed(B * H)
    PHn = (H + 1) // 2
    planes = torch.randn((B, PHn, PHn, 4, Cc), generator=g, device="cuda").to(torch.bfloat16)
    y = torch.relu(torch.randn((B, H, H, Cc), generator=g, device="cuda")).to(torch.bfloat16)
    dz = torch.empty((B, H, H, Cc), dtype=torch.bfloat16, device="cuda")
    db = torch.empty((Cc,), device="cuda")
    scratch = torch.zeros(int(L.ta_relu_bwd_bias_scratch_floats(B * H * H, Cc)), device="cuda")
    check(L.ta_planes_relu_bwd_bias(_p(planes), _p(y), _p(dz), B, H, H, Cc, _p(db), _p(scratch), _st()), "ta_planes_relu_bwd_bias")
    hh = torch.arange(H, device="cuda")
    dense = planes[:, hh[:, None] // 2, hh[None, :] // 2, (hh[:, None] % 2) * 2 + (hh[None, :] % 2)]      # [B,H,H,C]
    want = torch.where(y > 0, dense, torch.zeros_like(dense))
    assert torch.equal(dz, want)
    ws = want.double().sum((0, 1, 2))
    assert float((db.double() - ws).abs().max()) <= 1e-5 * max(1.0, float(want.double().abs().sum((0, 1, 2)).max()))


@pytest.mark.parametrize("B", [1, 7, 300, 1500])
@pytest.mark.parametrize("with_mask", [False, True])
@pytest.mark.parametrize("class_major", [1, 0])
def test_conv2_dgrad_planes_tcgen05(B, with_mask, class_major):
    """ta_conv2_dgrad_planes (tcgen05, per-class tap lists; class_major = 1 the warp-specialised kernel with bulk stores,
    0 the single-role kernel) == conv_transpose2d(dz, w, stride 2) regrouped into merged
    parity planes [B][17][17][4][64], evaluated in float64 from the same bf16 inputs: every entry within bf16 rounding of
    the exact value (2^-8 relative + 1e-3 of the tensor's scale for the fp32 accumulation of 256 products); entries whose
    pixel does not exist (row / column 33) are zero; with the first layer's ReLU bit mask the masked entries are zero.
    B = 1500 gives every persistent CTA several tiles (double-buffered cp.async staging, TMEM reuse)."""
    import torch.nn.functional as F
    L, check = _lib()
    g = torch.Generator(device="cuda").manual_seed(B)
    w = (torch.randn((64, 64, 3, 3), generator=g, device="cuda") * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dz = torch.randn((B, 16, 16, 64), generator=g, device="cuda").to(torch.bfloat16)
    wimg = torch.empty(9 * 64 * 64, dtype=torch.bfloat16, device="cuda")
    check(L.ta_conv2_dgrad_prep(_p(w), w.stride(0), w.stride(1), w.stride(2), w.stride(3), _p(wimg), _st()), "ta_conv2_dgrad_prep")
    mask = None
    if with_mask:
        mask = torch.randint(0, 2 ** 31 - 1, (B * 289 * 8,), generator=g, device="cuda", dtype=torch.int32)
        mask = mask ^ (torch.randint(0, 2, (B * 289 * 8,), generator=g, device="cuda", dtype=torch.int32) << 31)
    planes = torch.full((B, 17, 17, 4, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    guard = torch.zeros(4096, dtype=torch.bfloat16, device="cuda")      # allocated right behind: a write past the end would show
    check(L.ta_conv2_dgrad_planes(_p(dz), _p(wimg), _p(mask), B, class_major, _p(planes), _st()), "ta_conv2_dgrad_planes")
    torch.cuda.synchronize()
    assert float(guard.float().abs().max()) == 0.0
    if class_major:      # [4][B*289][64] (the warp-specialised kernel) -> the position-major view the checks below use
        planes = planes.view(4, B, 17, 17, 64).permute(1, 2, 3, 0, 4).contiguous()
    assert L.ta_debug_conv1_tc_failed() == 0
    dx = F.conv_transpose2d(dz.permute(0, 3, 1, 2).double(), w.double(), stride=2)          # [B,64,33,33]
    want = torch.zeros((B, 17, 17, 4, 64), dtype=torch.float64, device="cuda")
    for c in range(4):
        pa, pb = c >> 1, c & 1
        sub = dx[:, :, pa::2, pb::2].permute(0, 2, 3, 1)
        want[:, :sub.shape[1], :sub.shape[2], c] = sub
    if with_mask:
        mb = mask.view(B, 17, 17, 4, 2).to(torch.int64) & 0xFFFFFFFF
        q = torch.arange(16, device="cuda")
        even = ((mb[..., None] >> q) & 1).bool()                                             # [.., 2, 16]: channel half*32 + 2q
        odd = ((mb[..., None] >> (16 + q)) & 1).bool()
        keep = torch.stack([even, odd], -1).reshape(B, 17, 17, 4, 64)
        want = want * keep
    got = planes.double()
    assert bool(torch.isfinite(got).all())
    scale = float(want.abs().max())
    err = (got - want).abs()
    assert bool((err <= want.abs() * 2.0 ** -8 + 1e-3 * scale).all()), float((err - want.abs() * 2.0 ** -8).max() / scale)
    assert float(got[:, 16, :, 2:].abs().max()) == 0.0 and float(got[:, :, 16, 1::2].abs().max()) == 0.0      # pixels of row / column 33


@pytest.mark.parametrize("B", [1, 7, 300, 1500])
@pytest.mark.parametrize("codes", [True, False])
@pytest.mark.parametrize("with_mask", [True, False])
def test_conv1_fwd_ws_equals_single_role_kernel(B, codes, with_mask):
    """conv1_fwd_ws_kernel (warp-specialised, seven-grid-row tiles, one N = 256 MMA per operand pair, tensor-map stores of
    whole image rows) performs the same arithmetic as conv1_fwd_tc_kernel: outputs and ReLU bit masks are BIT-IDENTICAL, for
    code and float inputs, batch sizes with partial last tiles, and nothing is written outside y (NaN-filled guard rows)."""
    L, check = _lib()
    g = torch.Generator(device="cuda").manual_seed(B + 11)
    w4 = torch.randn((256, 16), generator=g, device="cuda") * 0.3
    b4 = torch.randn((256,), generator=g, device="cuda") * 0.1
    if codes:
        x = torch.tensor([0, 1, 2, 4], dtype=torch.uint8, device="cuda")[torch.randint(0, 4, (B, 5, 289), generator=g, device="cuda")]
        x_dtype = 1
    else:
        x, x_dtype = torch.randn((B, 5, 289), generator=g, device="cuda"), 0
    outs = []
    for mode in (2, 1):
        prev = L.ta_debug_conv1_tc(mode)
        try:
            y = torch.full((B + 1, 33, 33, 64), float("nan"), dtype=torch.bfloat16, device="cuda")    # the last sample is a guard
            mask = torch.full((B * 289 * 8 + 64,), -1, dtype=torch.int32, device="cuda")
            if with_mask:
                check(L.ta_conv1_fwd_mask(_p(x), x_dtype, x.stride(0), _p(w4), _p(b4), B, _p(y), _p(mask), _st()), "ta_conv1_fwd_mask")
            else:
                check(L.ta_conv1_fwd(_p(x), x_dtype, x.stride(0), _p(w4), _p(b4), B, _p(y), _st()), "ta_conv1_fwd")
            torch.cuda.synchronize()
        finally:
            L.ta_debug_conv1_tc(prev)
        assert L.ta_debug_conv1_tc_failed() == 0
        assert bool(torch.isnan(y[B].float()).all()) and bool((mask[B * 289 * 8:] == -1).all())
        assert bool(torch.isfinite(y[:B].float()).all())
        outs.append((y[:B].view(torch.int16).clone(), mask[:B * 289 * 8].clone()))
    assert bool((outs[0][0] == outs[1][0]).all())
    if with_mask:
        assert bool((outs[0][1] == outs[1][1]).all())


@pytest.mark.parametrize("B", [1, 7, 300, 1500])
@pytest.mark.parametrize("codes", [True, False])
def test_conv2_dgrad_conv1_bwd_fused(B, codes):
    """ta_conv2_dgrad_conv1_bwd (one kernel: conv2's data gradient on tcgen05 -> ReLU mask -> conv1's weight / bias gradient
    GEMM, the planes never leaving the SM) against an independent float64 evaluation from the same bf16 inputs:
        planes = bf16(conv_transpose2d(dz, w2, stride 2)) regrouped by output parity, masked with the bit mask,
        dW4[(c, ch), k] = sum over positions planes[pos, c, ch] * patch[pos, k],  db4 = sum over positions planes
    (patch = the decoded 2x2 input patch in the folded layer's tap order, conv1.fold).  Tolerance 2e-3 of the largest entry
    (the kernel rounds the planes to bf16 exactly like the two-kernel path; fp32 accumulation over up to 433 k positions),
    and agreement with the two-kernel path (ta_conv2_dgrad_planes + ta_conv1_bwd_planes) to 1e-3 of the largest entry."""
    import torch.nn.functional as F
    L, check = _lib()
    g = torch.Generator(device="cuda").manual_seed(B + 5)
    w = (torch.randn((64, 64, 3, 3), generator=g, device="cuda") * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dz = torch.randn((B, 16, 16, 64), generator=g, device="cuda").to(torch.bfloat16)
    wimg = torch.empty(9 * 64 * 64, dtype=torch.bfloat16, device="cuda")
    check(L.ta_conv2_dgrad_prep(_p(w), w.stride(0), w.stride(1), w.stride(2), w.stride(3), _p(wimg), _st()), "ta_conv2_dgrad_prep")
    mask = torch.randint(0, 2 ** 31 - 1, (B * 289 * 8,), generator=g, device="cuda", dtype=torch.int32)
    mask = mask ^ (torch.randint(0, 2, (B * 289 * 8,), generator=g, device="cuda", dtype=torch.int32) << 31)
    lut = torch.tensor([0.9, -0.9, -0.5, 0.0, 0.3], device="cuda")
    xc = torch.tensor([0, 1, 2, 4], dtype=torch.uint8, device="cuda")[torch.randint(0, 4, (B, 5, 289), generator=g, device="cuda")]
    if codes:
        x, x_dtype = xc, 1
    else:
        x, x_dtype = torch.randn((B, 5, 289), generator=g, device="cuda"), 0
    dw4 = torch.full((256, 16), float("nan"), device="cuda")
    db4 = torch.full((256,), float("nan"), device="cuda")
    check(L.ta_conv2_dgrad_conv1_bwd(_p(dz), _p(wimg), _p(mask), _p(x), x_dtype, x.stride(0), B, _p(dw4), _p(db4), _st()), "ta_conv2_dgrad_conv1_bwd")
    torch.cuda.synchronize()
    assert L.ta_debug_conv1_tc_failed() == 0
    # ---- the two-kernel path
    planes = torch.empty((4, B * 289, 64), dtype=torch.bfloat16, device="cuda")
    dw4b, db4b = torch.empty_like(dw4), torch.empty_like(db4)
    check(L.ta_conv2_dgrad_planes(_p(dz), _p(wimg), None, B, 1, _p(planes), _st()), "ta_conv2_dgrad_planes")
    check(L.ta_conv1_bwd_planes(_p(x), x_dtype, x.stride(0), None, _p(mask), _p(planes), 1, B, _p(dw4b), _p(db4b), _st()), "ta_conv1_bwd_planes")
    # ---- float64
    dx = F.conv_transpose2d(dz.permute(0, 3, 1, 2).double(), w.double(), stride=2)          # [B,64,33,33]
    pl = torch.zeros((B, 17, 17, 4, 64), dtype=torch.float64, device="cuda")
    for c in range(4):
        pa, pb = c >> 1, c & 1
        sub = dx[:, :, pa::2, pb::2].permute(0, 2, 3, 1)
        pl[:, :sub.shape[1], :sub.shape[2], c] = sub
    pl = pl.to(torch.bfloat16).double()
    mb = mask.view(B, 17, 17, 4, 2).to(torch.int64) & 0xFFFFFFFF
    q = torch.arange(16, device="cuda")
    even = ((mb[..., None] >> q) & 1).bool()
    odd = ((mb[..., None] >> (16 + q)) & 1).bool()
    pl = pl * torch.stack([even, odd], -1).reshape(B, 17, 17, 4, 64)
    vals = (lut[x[:, :4].long()] if codes else x[:, :4]).double().view(B, 4, 17, 17)          # frames 0..3 as a [17][17] grid (cell = m*17+n)
    vp = F.pad(vals, (0, 1, 0, 1))                                                          # neighbours past the edge are zero
    # tap k = (dm*2 + dn) * 4 + frame: the patch order of conv1.fold / the kernels' (d00, d01, d10, d11) x 4 frames
    patch = torch.stack([vp[:, f, dm:dm + 17, dn:dn + 17] for dm in (0, 1) for dn in (0, 1) for f in range(4)], -1)   # [B,17,17,16]
    want_w = torch.einsum("bmncd,bmnk->cdk", pl, patch).reshape(256, 16)
    want_b = pl.sum((0, 1, 2)).reshape(256)
    sw, sb_ = float(want_w.abs().max()), float(want_b.abs().max())
    assert bool(torch.isfinite(dw4).all()) and bool(torch.isfinite(db4).all())
    assert float((dw4b.double() - want_w).abs().max()) <= 2e-3 * sw       # (the reference construction itself, via the two-kernel path)
    assert float((dw4.double() - want_w).abs().max()) <= 2e-3 * sw
    assert float((db4.double() - want_b).abs().max()) <= 2e-3 * sb_
    assert float((dw4 - dw4b).abs().max()) <= 1e-3 * sw and float((db4 - db4b).abs().max()) <= 1e-3 * sb_


@pytest.mark.parametrize("B", [4096, 300, 1])
def test_ppo_actor_loss_and_gradient(B):
    """ta_ppo_actor_loss == the reference's lines (PPO.py:124-132: Categorical(probs=softmax(logits)).entropy / log_prob,
    ratio, clipped surrogate, mean) evaluated with torch autograd in fp32: loss to 1e-5, gradient to bf16 rounding of its
    entries (2^-8 relative + 1e-7), bias gradient = column sums of the emitted gradient."""
    from torch.distributions import Categorical
    L, check = _lib()
    g = torch.Generator(device="cuda").manual_seed(B)
    lg = torch.zeros((B, 8), device="cuda")
    lg[:, :5] = torch.randn((B, 5), generator=g, device="cuda") * 2
    lg16 = lg.to(torch.bfloat16)
    act = torch.randint(0, 5, (B,), generator=g, device="cuda")
    old = torch.log(torch.rand(B, generator=g, device="cuda") * 0.6 + 0.05)
    adv = torch.randn(B, generator=g, device="cuda")
    clip, ent = 0.1, 0.01
    x = lg16[:, :5].float().clone().requires_grad_(True)
    dist = Categorical(probs=torch.softmax(x, 1))
    ratio = torch.exp(dist.log_prob(act) - old)
    loss = (-torch.min(ratio * adv, torch.clamp(ratio, 1 - clip, 1 + clip) * adv) - ent * dist.entropy()).mean()
    loss.backward()
    d = torch.empty((B, 8), dtype=torch.bfloat16, device="cuda")
    out = torch.zeros(1, device="cuda"); dbh = torch.zeros(5, device="cuda"); step = torch.full((1,), 4.0, device="cuda")
    check(L.ta_ppo_actor_loss(_p(lg16), _p(act.to(torch.int32)), _p(old), _p(adv), B, clip, ent, _p(d), _p(out), _p(dbh), _p(step), _st()),
          "ta_ppo_actor_loss")
    assert float(out) == pytest.approx(float(loss), rel=1e-5, abs=1e-6)
    got = d[:, :5].float()
    assert bool(((got - x.grad).abs() <= x.grad.abs() * 2.0 ** -8 + 1e-7).all())
    assert float(d[:, 5:].float().abs().max()) == 0.0
    assert torch.allclose(dbh, got.sum(0), rtol=1e-5, atol=1e-7)
    assert float(step) == 5.0


@pytest.mark.parametrize("B", [4096, 77])
def test_ppo_critic_loss_and_gradient(B):
    L, check = _lib()
    g = torch.Generator(device="cuda").manual_seed(B)
    v = torch.zeros((B, 8), device="cuda")
    v[:, 0] = torch.randn(B, generator=g, device="cuda") * 1.5
    v16 = v.to(torch.bfloat16)
    tgt = torch.randn(B, generator=g, device="cuda")
    x = v16[:, 0].float().clone().requires_grad_(True)
    loss = torch.nn.functional.smooth_l1_loss(x.view(-1, 1), tgt.view(-1, 1))
    loss.backward()
    d = torch.empty((B, 8), dtype=torch.bfloat16, device="cuda")
    out = torch.zeros(1, device="cuda"); dbh = torch.zeros(1, device="cuda")
    check(L.ta_ppo_critic_loss(_p(v16), _p(tgt), B, _p(d), _p(out), _p(dbh), None, _st()), "ta_ppo_critic_loss")
    assert float(out) == pytest.approx(float(loss), rel=1e-5, abs=1e-7)
    assert bool(((d[:, 0].float() - x.grad).abs() <= x.grad.abs() * 2.0 ** -8 + 1e-9).all())
    assert float(dbh) == pytest.approx(float(d[:, 0].float().sum()), rel=1e-5, abs=1e-8)


def test_adam_shadow_matches_torch_adam():
    """ta_adam_shadow == torch.optim.Adam(lr=1e-4, eps=1e-5) over 5 steps on a flat buffer (1e-6 relative on the parameter
    after each step), bf16 shadow = the rounded parameter, gradient scale folded in."""
    L, check = _lib()
    n = 100003
    g = torch.Generator(device="cuda").manual_seed(1)
    p0 = torch.randn(n, generator=g, device="cuda")
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-4, eps=1e-5)
    p = p0.clone(); m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
    p16 = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    step = torch.zeros(1, device="cuda")
    for t in range(5):
        grad = torch.randn(n, generator=g, device="cuda") * (10.0 ** (t - 3))
        ref.grad = grad.clone()
        opt.step()
        step += 1
        check(L.ta_adam_shadow(_p(p), _p((grad * 4).contiguous()), _p(m), _p(v), _p(p16), n, _p(step), 1e-4, 0.9, 0.999, 1e-5, 0.25, _st()),
              "ta_adam_shadow")
        assert float((p - ref.detach()).abs().max()) <= 1e-6 * float(ref.detach().abs().max()) + 2e-7 * (t + 1)
        assert torch.equal(p16, p.to(torch.bfloat16))
    st = opt.state[ref]
    assert torch.allclose(m, st["exp_avg"], rtol=1e-5, atol=1e-9) and torch.allclose(v, st["exp_avg_sq"], rtol=1e-5, atol=1e-12)


def test_weight_prep_and_gradient_finalise_are_adjoint_forms_of_the_module_weights():
    """ta_tinet_prep == conv1.fold (the einsum the autograd path uses) + the fc0 permutation + zero padding; ta_tinet_grad
    maps gradients in those forms back exactly as autograd does through fold / the permuted view / the slices."""
    import twoarmy_b200 as pkg
    P = importlib.import_module(pkg.__name__ + ".ppo")
    FS = importlib.import_module(pkg.__name__ + ".fused_step")
    C1 = importlib.import_module(pkg.__name__ + ".conv1")
    torch.manual_seed(0)
    for kind, net in (("actor", P.Net_PPO_actor()), ("critic", P.Net_PPO_critic())):
        net = net.cuda().to(memory_format=torch.channels_last)
        ref = {k: v.detach().clone() for k, v in net.state_dict().items()}
        fn = FS.FusedNet(net, kind, 1e-4, 1e-5)
        for k, v in net.state_dict().items():          # re-homing the parameters into the flat buffer changes no value
            assert torch.equal(v, ref[k]), k
        bone = net.bone1 if kind == "actor" else net.bone2
        head = net.A if kind == "actor" else net.V
        w4, b4 = C1.fold(bone.cnn_base[0].weight.detach(), bone.cnn_base[0].bias.detach())
        assert torch.allclose(fn.w4, w4, rtol=1e-6, atol=1e-7) and torch.equal(fn.b4, b4)
        w0 = bone.fc0.weight.detach().view(256, 256, 9).permute(0, 2, 1).reshape(256, 2304).to(torch.bfloat16)
        assert torch.equal(fn.fc0p, w0)
        assert torch.equal(fn.pos16[:, :10], bone.positionnet.weight.detach().to(torch.bfloat16)) and float(fn.pos16[:, 10:].float().abs().max()) == 0
        nh = head.weight.shape[0]
        assert torch.equal(fn.head8[:nh], head.weight.detach().to(torch.bfloat16)) and float(fn.head8[nh:].float().abs().max()) == 0
        assert torch.equal(fn.p16["w2"], bone.cnn_base[2].weight.detach().to(torch.bfloat16))
        # gradients: random staged gradients -> flat buffer, against autograd through the same forms
        g = torch.Generator(device="cuda").manual_seed(3)
        fn.dw4.copy_(torch.randn((256, 16), generator=g, device="cuda")); fn.db4.copy_(torch.randn(256, generator=g, device="cuda"))
        for t in (fn.g_fc0p, fn.g_pos16, fn.g_head8, fn.g_w4c, fn.g_wfc1):
            t.copy_(torch.randn(t.shape, generator=g, device="cuda").to(torch.bfloat16))
        gw2 = torch.randn((64, 64, 3, 3), generator=g, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        gw3 = torch.randn((128, 64, 4, 4), generator=g, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        a = FS._GradArgs()
        a.dw4, a.db4 = fn.dw4.data_ptr(), fn.db4.data_ptr()
        a.g_w1, a.g_b1 = fn.g32["w1"].data_ptr(), fn.g32["b1"].data_ptr()
        a.s_o, a.s_c, a.s_y, a.s_x = fn.g32["w1"].stride()
        srcs, dsts = (gw2, gw3, fn.g_w4c, fn.g_wfc1), (fn.g32["w2"], fn.g32["w3"], fn.g32["w4c"], fn.g32["wfc1"])
        for k in range(4):
            a.src[k], a.dst[k], a.n[k] = srcs[k].data_ptr(), dsts[k].data_ptr(), dsts[k].numel()
        a.fc0p, a.g_fc0 = fn.g_fc0p.data_ptr(), fn.g32["wfc0"].data_ptr()
        a.pos16, a.g_pos = fn.g_pos16.data_ptr(), fn.g32["wpos"].data_ptr()
        a.head8, a.g_head = fn.g_head8.data_ptr(), fn.g32["wh"].data_ptr()
        a.nh = nh
        L, check = _lib()
        check(L.ta_tinet_grad(C.byref(a), _st()), "ta_tinet_grad")
        w1 = bone.cnn_base[0].weight.detach().clone().requires_grad_(True)
        b1 = bone.cnn_base[0].bias.detach().clone().requires_grad_(True)
        w4r, b4r = C1.fold(w1, b1)
        ((w4r * fn.dw4).sum() + (b4r * fn.db4).sum()).backward()
        assert torch.allclose(bone.cnn_base[0].weight.grad, w1.grad, rtol=1e-5, atol=1e-6)
        assert torch.allclose(bone.cnn_base[0].bias.grad, b1.grad, rtol=1e-5, atol=1e-6)
        assert torch.equal(bone.cnn_base[2].weight.grad, gw2.float()) and torch.equal(bone.cnn_base[4].weight.grad, gw3.float())
        assert torch.equal(bone.cnn_base[6].weight.grad.permute(0, 2, 3, 1).reshape(256, 1152), fn.g_w4c.float())
        assert torch.equal(bone.fc1.weight.grad, fn.g_wfc1.float())
        assert torch.equal(bone.fc0.weight.grad.view(256, 256, 9).permute(0, 2, 1).reshape(256, 2304), fn.g_fc0p.float())
        assert torch.equal(bone.positionnet.weight.grad, fn.g_pos16[:, :10].float())
        assert torch.equal(head.weight.grad, fn.g_head8[:nh].float())


def test_gather_minibatch():
    L, check = _lib()
    g = torch.Generator(device="cuda").manual_seed(5)
    N, bs = 1000, 333
    s = torch.randint(0, 5, (N, 5, 289), generator=g, device="cuda", dtype=torch.uint8)
    p = torch.randint(1, 16, (N, 5, 2), generator=g, device="cuda").float()
    M = 1400   # samples: the N records followed by 400 relabelled copies
    src = torch.cat([torch.arange(N, device="cuda"), torch.randint(0, N, (M - N,), generator=g, device="cuda")])
    gg = torch.randint(1, 16, (M, 2), generator=g, device="cuda").float()
    a = torch.randint(0, 5, (M, 1), generator=g, device="cuda")
    old = torch.randn((M, 1), generator=g, device="cuda"); adv = torch.randn((M, 1), generator=g, device="cuda"); tv = torch.randn((M, 1), generator=g, device="cuda")
    idx = torch.randperm(M, generator=g, device="cuda")[:bs]
    for use_src in (True, False):
        sb = torch.empty((bs, 4, 289), dtype=torch.uint8, device="cuda"); pg = torch.empty((bs, 16), dtype=torch.bfloat16, device="cuda")
        a_mb = torch.empty(bs, dtype=torch.int32, device="cuda"); o_mb = torch.empty(bs, device="cuda"); ad_mb = torch.empty(bs, device="cuda"); tv_mb = torch.empty(bs, device="cuda")
        ii = idx if use_src else idx % N
        check(L.ta_gather_minibatch(_p(s), _p(p), _p(gg), _p(a), _p(old), _p(adv), _p(tv), _p(ii), _p(src) if use_src else None, bs, _p(sb), _p(pg),
                                    _p(a_mb), _p(o_mb), _p(ad_mb), _p(tv_mb), _st()), "ta_gather_minibatch")
        rec = src[ii] if use_src else ii
        assert torch.equal(sb, s[rec][:, 0:4])
        assert torch.equal(pg[:, :8].float(), p[rec][:, 0:4].reshape(bs, 8)) and torch.equal(pg[:, 8:10].float(), gg[ii])
        assert float(pg[:, 10:].float().abs().max()) == 0
        assert torch.equal(a_mb.long(), a[ii, 0]) and torch.equal(o_mb, old[ii, 0]) and torch.equal(ad_mb, adv[ii, 0]) and torch.equal(tv_mb, tv[ii, 0])
