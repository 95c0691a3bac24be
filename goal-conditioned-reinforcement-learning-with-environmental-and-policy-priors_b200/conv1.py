"""TINet's first layer through the fused CUDA kernels (ta_conv1_fwd / ta_conv1_bwd).

UpsamplingNearest2d(4) + Conv2d(4, 64, kernel 4, stride 2) + ReLU (soa/agent/net/all_net.py:142-143,
157,180-181) applied to 17x17 frames is folded exactly: output pixel (2m+py, 2n+px) sees input pixels
(m+dy, n+dx) through phase-dependent sums of kernel taps.  `fold` builds that [256,16] weight with an
einsum (differentiable, so the conv's own parameters receive the gradient), the kernels do the K = 16
products, bias, ReLU and the bf16 channels-last store, straight from the rollout buffer's uint8 codes or
from float frames.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _capi

# Sel[p][d][k]: kernel tap k of one axis lands on input offset d for output phase p
_SEL = (((1., 1., 1., 1.), (0., 0., 0., 0.)), ((1., 1., 0., 0.), (0., 0., 1., 1.)))


_sel_cache = {}


def fold(weight: torch.Tensor, bias: torch.Tensor):
    """conv weight [64,4,4,4], bias [64] -> (w4 [256,16] rows (py,px,o) cols (dy,dx,c), b4 [256]), float32."""
    key = str(weight.device)
    if key not in _sel_cache:  # built once per device (a host-to-device copy is not capturable in a CUDA graph)
        _sel_cache[key] = torch.tensor(_SEL, dtype=torch.float32, device=weight.device)
    sel = _sel_cache[key]
    w4 = torch.einsum("ocyx,pdy,qex->pqodec", weight.float(), sel, sel).reshape(256, 16)
    return w4.contiguous(), bias.float().repeat(4).contiguous()


def _ptr(t):
    return C.c_void_p(t.data_ptr())


class _Conv1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w4, b4):
        # x [B, >=4, 289] uint8 codes or float32; only frames 0..3 of each sample are read
        assert x.is_cuda and x.dim() == 3 and x.shape[2] == 289 and x.shape[1] >= 4
        if x.dtype not in (torch.uint8, torch.float32):
            x = x.float()
        if x.stride(2) != 1 or x.stride(1) != 289 or x.stride(0) < 4 * 289:  # e.g. an expanded constant stack
            x = x.contiguous()
        B = x.shape[0]
        y = torch.empty((B, 33, 33, 64), dtype=torch.bfloat16, device=x.device)
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        w4, b4 = w4.detach().float().contiguous(), b4.detach().float().contiguous()
        _capi.check(_capi.lib().ta_conv1_fwd(_ptr(x), 1 if x.dtype == torch.uint8 else 0, x.stride(0), _ptr(w4), _ptr(b4), B,
                                             _ptr(y), st), "ta_conv1_fwd")
        ctx.save_for_backward(x, y)
        return y.permute(0, 3, 1, 2)  # logical [B,64,33,33] in channels-last memory

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        dy = dy.permute(0, 2, 3, 1).to(torch.bfloat16).contiguous()
        dw4 = torch.empty((256, 16), dtype=torch.float32, device=x.device)
        db4 = torch.empty((256,), dtype=torch.float32, device=x.device)
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _capi.check(_capi.lib().ta_conv1_bwd(_ptr(x), 1 if x.dtype == torch.uint8 else 0, x.stride(0), _ptr(y), _ptr(dy),
                                             x.shape[0], _ptr(dw4), _ptr(db4), st), "ta_conv1_bwd")
        return None, dw4, db4


def conv1_relu(x: torch.Tensor, conv: torch.nn.Conv2d) -> torch.Tensor:
    """relu(conv(upsample4(decode(x)))) as bf16 [B,64,33,33] (channels-last)."""
    w4, b4 = fold(conv.weight, conv.bias)
    return _Conv1.apply(x, w4, b4)


class _ConvS2(torch.autograd.Function):
    """relu(Conv2d(stride 2, no padding)) on channels-last bf16.  Forward: cuDNN's fused conv + bias + ReLU
    (111 us instead of 314 us for conv2 at B = 4096: no separate bias-add / ReLU passes).  Backward: weight
    gradient from cuDNN; bias gradient from the channel-sum kernel (29 us instead of aten::sum's 112 us);
    DATA gradient, for which cuDNN's strided-dgrad kernels take 1-1.7 ms, as a plain cuBLAS GEMM (dZ x W)
    followed by the col2im gather kernel (ta_col2im_s2)."""

    @staticmethod
    def forward(ctx, x, w, b):
        y = torch.cudnn_convolution_relu(x, w, b, [2, 2], [0, 0], [1, 1], 1)
        ctx.save_for_backward(x, w, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dz = torch.ops.aten.threshold_backward(dy, y, 0).contiguous(memory_format=torch.channels_last)
        cout, cin, k, _ = w.shape
        lib, st = _capi.lib(), C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _, gw, _ = torch.ops.aten.convolution_backward(dz, x, w, None, [2, 2], [0, 0], [1, 1], False, [0, 0], 1,
                                                       [False, True, False])
        dz_rows = dz.permute(0, 2, 3, 1).reshape(-1, cout)            # [B*OH*OW, cout], a view of channels-last memory
        gb32 = torch.empty((cout,), dtype=torch.float32, device=x.device)
        _capi.check(lib.ta_channel_sum_bf16(_ptr(dz_rows), dz_rows.shape[0], cout, _ptr(gb32), st), "ta_channel_sum_bf16")
        gx = None
        if ctx.needs_input_grad[0]:
            B, _, H, W = x.shape
            dcols = dz_rows @ w.permute(0, 2, 3, 1).reshape(cout, k * k * cin)
            gx_nhwc = torch.empty((B, H, W, cin), dtype=torch.bfloat16, device=x.device)
            _capi.check(lib.ta_col2im_s2(_ptr(dcols), _ptr(gx_nhwc), B, H, W, cin, k, st), "ta_col2im_s2")
            gx = gx_nhwc.permute(0, 3, 1, 2)
        return gx, gw, gb32.to(w.dtype)


def conv_s2_relu(x: torch.Tensor, conv: torch.nn.Conv2d) -> torch.Tensor:
    """relu(conv(x)) for a stride-2 unpadded Conv2d, x bf16 channels-last (see _ConvS2)."""
    return _ConvS2.apply(x.contiguous(memory_format=torch.channels_last),
                         conv.weight.to(torch.bfloat16).contiguous(memory_format=torch.channels_last),
                         conv.bias.to(torch.bfloat16))


def parity_class_weights(w: torch.Tensor) -> torch.Tensor:
    """conv weight [cout, cin, k, k] (stride 2, no padding, k = 3 or 4) -> the weight [4*cin, cout, 2, 2] (channels-last)
    of the ONE stride-1 convolution (padding 1) that produces its data gradient as merged parity planes: output
    channel block c = pa*2+pb at position (i, j) is dx[2i+pa, 2j+pb], which only receives the taps ky = pa, kx = pb
    (mod 2) from dz[i - ky // 2, j - kx // 2].  A parity with a single tap along an axis (k = 3, odd parity) has a
    zero in its other kernel slot, and its extra last row / column comes out as zero."""
    cout, cin, k, _ = w.shape
    if w.is_cuda and w.dtype == torch.bfloat16:  # one kernel instead of two dozen slice / flip / pad / copy launches
        buf = torch.empty((4 * cin, 2, 2, cout), dtype=torch.bfloat16, device=w.device)
        st = C.c_void_p(torch.cuda.current_stream(w.device).cuda_stream)
        _capi.check(_capi.lib().ta_parity_class_weights(_ptr(w), w.stride(0), w.stride(1), w.stride(2), w.stride(3), cout, cin, k,
                                                        _ptr(buf), st), "ta_parity_class_weights")
        return buf.permute(0, 3, 1, 2)
    blocks = []
    for pa in (0, 1):
        for pb in (0, 1):
            sub = w[:, :, pa::2][:, :, :, pb::2].flip(2, 3)               # taps {0, 2} / {1, 3} (k = 4) or {1} for parity 0 / 1
            sub = torch.nn.functional.pad(sub, (2 - sub.shape[3], 0, 2 - sub.shape[2], 0))   # single tap -> slot 1
            blocks.append(sub.permute(1, 0, 2, 3))
    return torch.cat(blocks, 0).contiguous(memory_format=torch.channels_last)


def _class_planes(dz: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """The data gradient of a k x k stride-2 unpadded conv with weight w, given dz (channels-last bf16), as merged
    parity planes: contiguous [B, OH+1, OW+1, 4*cin] (see parity_class_weights)."""
    pl = torch.nn.functional.conv2d(dz, parity_class_weights(w), padding=1)
    return pl.permute(0, 2, 3, 1).contiguous()


def _wgrad_bgrad(dz: torch.Tensor, x_nchw: torch.Tensor, w: torch.Tensor, st):
    """cuDNN weight gradient + channel-sum bias gradient of a stride-2 unpadded conv (dz, x channels-last bf16)."""
    cout = w.shape[0]
    _, gw, _ = torch.ops.aten.convolution_backward(dz, x_nchw, w, None, [2, 2], [0, 0], [1, 1], False, [0, 0], 1, [False, True, False])
    dz_rows = dz.permute(0, 2, 3, 1).reshape(-1, cout)
    gb = torch.empty((cout,), dtype=torch.float32, device=dz.device)
    _capi.check(_capi.lib().ta_channel_sum_bf16(_ptr(dz_rows), dz_rows.shape[0], cout, _ptr(gb), st), "ta_channel_sum_bf16")
    return gw, gb.to(w.dtype)


class _Stem(torch.autograd.Function):
    """conv1 (fused kernel) + conv2 + conv3 (cuDNN fused conv + bias + ReLU) as ONE autograd node, so that the data
    gradients of the two stride-2 convolutions can stay in the form that is cheapest to produce: ONE stride-1 cuDNN
    convolution of dz whose output channels hold the four parity classes of the input pixel ("merged planes"; several
    times faster than the plain GEMM + col2im path or cuDNN's strided dgrad, scripts/probe_dgrad_classes.py).
      * conv3's planes are interleaved into the dense dz2 by the kernel that applies conv2's ReLU mask
        (ta_planes_to_dense_relu: the traffic of the threshold_backward pass it replaces);
      * conv2's planes are exactly the four output phases of the folded conv1, whose tcgen05 weight-gradient
        kernel reads them in place (ta_conv1_bwd_planes): the 33x33x64 gradient tensor is never materialised."""

    @staticmethod
    def forward(ctx, x, w4, b4, w2, b2, w3, b3):
        assert x.is_cuda and x.dim() == 3 and x.shape[2] == 289 and x.shape[1] >= 4
        if x.dtype not in (torch.uint8, torch.float32):
            x = x.float()
        if x.stride(2) != 1 or x.stride(1) != 289 or x.stride(0) < 4 * 289:
            x = x.contiguous()
        B = x.shape[0]
        y1 = torch.empty((B, 33, 33, 64), dtype=torch.bfloat16, device=x.device)
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        w4d, b4d = w4.detach().float().contiguous(), b4.detach().float().contiguous()
        # the layer's ReLU mask as bits (8 bytes per output pixel): what the weight-gradient kernel reads instead of y1
        mask = torch.empty((B * 289 * 8,), dtype=torch.int32, device=x.device)
        _capi.check(_capi.lib().ta_conv1_fwd_mask(_ptr(x), 1 if x.dtype == torch.uint8 else 0, x.stride(0), _ptr(w4d), _ptr(b4d), B,
                                                  _ptr(y1), _ptr(mask), st), "ta_conv1_fwd_mask")
        y2 = torch.cudnn_convolution_relu(y1.permute(0, 3, 1, 2), w2, b2, [2, 2], [0, 0], [1, 1], 1)
        if w3 is None:
            ctx.save_for_backward(x, y1, w2, y2, mask)
            return y2
        y3 = torch.cudnn_convolution_relu(y2, w3, b3, [2, 2], [0, 0], [1, 1], 1)
        ctx.save_for_backward(x, y1, w2, y2, mask, w3, y3)
        return y3

    @staticmethod
    def backward(ctx, dy):
        saved = ctx.saved_tensors
        x, y1, w2, y2, mask = saved[:5]
        lib, st = _capi.lib(), C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        gw3 = gb3 = None
        if len(saved) == 7:
            w3, y3 = saved[5:]
            dz3 = torch.ops.aten.threshold_backward(dy, y3, 0).contiguous(memory_format=torch.channels_last)
            gw3, gb3 = _wgrad_bgrad(dz3, y2, w3, st)
            p3 = _class_planes(dz3, w3)
            Bn, Cc, H, W = y2.shape
            dz2_nhwc = torch.empty((Bn, H, W, Cc), dtype=torch.bfloat16, device=x.device)
            y2_nhwc = y2.permute(0, 2, 3, 1)
            assert y2_nhwc.is_contiguous()
            _capi.check(lib.ta_planes_to_dense_relu(_ptr(p3), _ptr(y2_nhwc), _ptr(dz2_nhwc), Bn, H, W, Cc, w3.shape[2], st),
                        "ta_planes_to_dense_relu")
            dz2 = dz2_nhwc.permute(0, 3, 1, 2)
        else:
            dz2 = torch.ops.aten.threshold_backward(dy, y2, 0).contiguous(memory_format=torch.channels_last)
        gw2, gb2 = _wgrad_bgrad(dz2, y1.permute(0, 3, 1, 2), w2, st)
        p2 = _class_planes(dz2, w2)
        dw4 = torch.empty((256, 16), dtype=torch.float32, device=x.device)
        db4 = torch.empty((256,), dtype=torch.float32, device=x.device)
        _capi.check(lib.ta_conv1_bwd_planes(_ptr(x), 1 if x.dtype == torch.uint8 else 0, x.stride(0), None, _ptr(mask), _ptr(p2), 0,
                                            x.shape[0], _ptr(dw4), _ptr(db4), st), "ta_conv1_bwd_planes")
        return None, dw4, db4, gw2, gb2, gw3, gb3


class _Stem8(torch.autograd.Function):
    """_Stem for a first convolution with EIGHT input channels (Net_PPO_Predictor_actor / _critic, all_net.py:249-305: the
    4 current frames + the 4 predicted ones).  The folded layer is linear in its input channels, so it is two passes of
    the 4-channel kernel: z_b = w4b . patch(x_b) (no bias, no ReLU), then y1 = relu(b4 + w4a . patch(x_a) + z_b)
    (ta_conv1_fwd_add); backward: conv2's data gradient as parity planes, read in place by the tcgen05 weight-gradient kernel
    once per half (ta_conv1_bwd_planes with y1 as the ReLU mask), the bias gradient from the first."""

    @staticmethod
    def forward(ctx, xa, xb, w4a, b4, w4b, w2, b2, w3, b3):
        xs = []
        for x in (xa, xb):
            assert x.is_cuda and x.dim() == 3 and x.shape[2] == 289 and x.shape[1] >= 4
            if x.dtype not in (torch.uint8, torch.float32):
                x = x.float()
            if x.stride(2) != 1 or x.stride(1) != 289 or x.stride(0) < 4 * 289:
                x = x.contiguous()
            xs.append(x)
        xa, xb = xs
        B = xa.shape[0]
        lib, st = _capi.lib(), C.c_void_p(torch.cuda.current_stream(xa.device).cuda_stream)
        zb = torch.empty((B, 33, 33, 64), dtype=torch.bfloat16, device=xa.device)
        y1 = torch.empty_like(zb)
        w4a_d, w4b_d, b4_d = (t.detach().float().contiguous() for t in (w4a, w4b, b4))
        zero = torch.zeros_like(b4_d)
        dt = lambda x: 1 if x.dtype == torch.uint8 else 0
        _capi.check(lib.ta_conv1_fwd_add(_ptr(xb), dt(xb), xb.stride(0), _ptr(w4b_d), _ptr(zero), B, None, 0, _ptr(zb), st), "ta_conv1_fwd_add")
        _capi.check(lib.ta_conv1_fwd_add(_ptr(xa), dt(xa), xa.stride(0), _ptr(w4a_d), _ptr(b4_d), B, _ptr(zb), 1, _ptr(y1), st), "ta_conv1_fwd_add")
        del zb
        y2 = torch.cudnn_convolution_relu(y1.permute(0, 3, 1, 2), w2, b2, [2, 2], [0, 0], [1, 1], 1)
        y3 = torch.cudnn_convolution_relu(y2, w3, b3, [2, 2], [0, 0], [1, 1], 1)
        ctx.save_for_backward(xa, xb, y1, w2, y2, w3, y3)
        return y3

    @staticmethod
    def backward(ctx, dy):
        xa, xb, y1, w2, y2, w3, y3 = ctx.saved_tensors
        lib, st = _capi.lib(), C.c_void_p(torch.cuda.current_stream(xa.device).cuda_stream)
        dz3 = torch.ops.aten.threshold_backward(dy, y3, 0).contiguous(memory_format=torch.channels_last)
        gw3, gb3 = _wgrad_bgrad(dz3, y2, w3, st)
        p3 = _class_planes(dz3, w3)
        Bn, Cc, H, W = y2.shape
        dz2_nhwc = torch.empty((Bn, H, W, Cc), dtype=torch.bfloat16, device=xa.device)
        y2_nhwc = y2.permute(0, 2, 3, 1)
        assert y2_nhwc.is_contiguous()
        _capi.check(lib.ta_planes_to_dense_relu(_ptr(p3), _ptr(y2_nhwc), _ptr(dz2_nhwc), Bn, H, W, Cc, w3.shape[2], st),
                    "ta_planes_to_dense_relu")
        dz2 = dz2_nhwc.permute(0, 3, 1, 2)
        gw2, gb2 = _wgrad_bgrad(dz2, y1.permute(0, 3, 1, 2), w2, st)
        p2 = _class_planes(dz2, w2)
        outs = []
        dt = lambda x: 1 if x.dtype == torch.uint8 else 0
        for x in (xa, xb):
            dw4 = torch.empty((256, 16), dtype=torch.float32, device=xa.device)
            db4 = torch.empty((256,), dtype=torch.float32, device=xa.device)
            _capi.check(lib.ta_conv1_bwd_planes(_ptr(x), dt(x), x.stride(0), _ptr(y1), None, _ptr(p2), 0, x.shape[0], _ptr(dw4), _ptr(db4), st),
                        "ta_conv1_bwd_planes")
            outs.append((dw4, db4))
        return None, None, outs[0][0], outs[0][1], outs[1][0], gw2, gb2, gw3, gb3


def stem8_relu(x8: torch.Tensor, conv1: torch.nn.Conv2d, conv2: torch.nn.Conv2d, conv3: torch.nn.Conv2d) -> torch.Tensor:
    """stem_relu for x8 [B, 8, 289] (float32: 4 current + 4 predicted frames) and a conv1 with 8 input channels."""
    w4a, b4 = fold(conv1.weight[:, 0:4], conv1.bias)
    w4b, _ = fold(conv1.weight[:, 4:8], conv1.bias)
    w2, b2 = _bf16_cl(conv2)
    w3, b3 = _bf16_cl(conv3)
    return _Stem8.apply(x8[:, 0:4], x8[:, 4:8], w4a, b4, w4b, w2, b2, w3, b3)


def _bf16_cl(conv: torch.nn.Conv2d):
    return conv.weight.to(torch.bfloat16).contiguous(memory_format=torch.channels_last), conv.bias.to(torch.bfloat16)


def stem_relu(x: torch.Tensor, conv1: torch.nn.Conv2d, conv2: torch.nn.Conv2d, conv3: torch.nn.Conv2d = None) -> torch.Tensor:
    """relu(conv3(relu(conv2(relu(conv1(upsample4(decode(x)))))))) as bf16 channels-last ([B,128,7,7]; without conv3
    [B,128,16,16]); see _Stem."""
    w4, b4 = fold(conv1.weight, conv1.bias)
    w2, b2 = _bf16_cl(conv2)
    w3, b3 = _bf16_cl(conv3) if conv3 is not None else (None, None)
    return _Stem.apply(x, w4, b4, w2, b2, w3, b3)


class _LinearReLU(torch.autograd.Function):
    """relu(x @ w.T + b) in bf16 with the bias + ReLU in the GEMM's epilogue (cuBLASLt) and the bias gradient from the
    channel-sum kernel: aten's column sum of a [4096, 512] matrix takes 23 us, latency-bound, eight times per
    optimiser step."""

    @staticmethod
    def forward(ctx, x, w, b):
        y = torch._addmm_activation(b, x, w.t(), use_gelu=False)
        ctx.save_for_backward(x, w, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dz = torch.ops.aten.threshold_backward(dy.contiguous(), y, 0)
        gb = torch.empty((w.shape[0],), dtype=torch.float32, device=x.device)
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _capi.check(_capi.lib().ta_channel_sum_bf16(_ptr(dz), dz.shape[0], dz.shape[1], _ptr(gb), st), "ta_channel_sum_bf16")
        gx = dz @ w if ctx.needs_input_grad[0] else None
        return gx, dz.t() @ x, gb.to(w.dtype)


def linear_relu(x: torch.Tensor, lin: torch.nn.Linear) -> torch.Tensor:
    """relu(lin(x)) for bf16 x on the GPU (out_features a power of two in 64..2048); see _LinearReLU."""
    return _LinearReLU.apply(x.contiguous(), lin.weight.to(torch.bfloat16), lin.bias.to(torch.bfloat16))


class _Im2colS2(torch.autograd.Function):
    """Patches of a channels-last bf16 map for a stride-2 unpadded k x k conv: forward the im2col kernel
    (ta_im2col_s2; index_select over pixels ran at a quarter of the HBM rate), backward the col2im kernel
    (instead of index_add_ with atomics).  `index` only documents the patch order (tests compare with it)."""

    @staticmethod
    def forward(ctx, x_nhwc, index, k):
        B, H, W, Cc = x_nhwc.shape
        ctx.shape, ctx.k = (B, H, W, Cc), k
        x_nhwc = x_nhwc.contiguous()
        OH, OW = (H - k) // 2 + 1, (W - k) // 2 + 1
        cols = torch.empty((B, OH * OW * k * k, Cc), dtype=torch.bfloat16, device=x_nhwc.device)
        st = C.c_void_p(torch.cuda.current_stream(x_nhwc.device).cuda_stream)
        _capi.check(_capi.lib().ta_im2col_s2(_ptr(x_nhwc), _ptr(cols), B, H, W, Cc, k, st), "ta_im2col_s2")
        return cols                                                        # [B, OH*OW*k*k, C]

    @staticmethod
    def backward(ctx, dcols):
        B, H, W, Cc = ctx.shape
        dcols = dcols.to(torch.bfloat16).contiguous()
        gx = torch.empty((B, H, W, Cc), dtype=torch.bfloat16, device=dcols.device)
        st = C.c_void_p(torch.cuda.current_stream(dcols.device).cuda_stream)
        _capi.check(_capi.lib().ta_col2im_s2(_ptr(dcols), _ptr(gx), B, H, W, Cc, ctx.k, st), "ta_col2im_s2")
        return gx, None, None


def im2col_s2(x_nhwc: torch.Tensor, index: torch.Tensor, k: int) -> torch.Tensor:
    return _Im2colS2.apply(x_nhwc, index, k)
