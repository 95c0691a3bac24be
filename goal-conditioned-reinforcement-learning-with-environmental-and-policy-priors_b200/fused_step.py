"""The PPO optimiser step, hand-scheduled: forward, backward and Adam of one TINet-based network as an explicit
list of kernels -- no autograd graph, no per-step weight casts, no gradient-accumulation kernels.

Mirrors soa/agent/PPO.py:124-144 (one minibatch: actor loss / critic loss, two backward passes, two Adam steps) on
soa/agent/net/all_net.py:139-247 (TINet + head).  What stays a library call is what is GEMM-shaped and large: the
second / third convolution and their weight gradients (cuDNN), the parity-plane data-gradient convolutions (cuDNN),
conv4 / Linear GEMMs (cuBLASLt, bias + ReLU in the epilogue).  Everything else is one of this repo's kernels:

    ta_gather_minibatch        sample gather (frames, positions + goal, action, old log-prob, advantage, target)
    ta_conv1_fwd_mask          LUT decode + Upsample(4) + conv1 + bias + ReLU on tcgen05, ReLU bit mask
    ta_im2col_s2 / ta_col2im_s2, ta_parity_class_weights, ta_planes_to_dense_relu, ta_conv1_bwd_planes
    ta_relu_bwd_bias           ReLU backward + bias gradient of every other layer, one pass, deterministic
    ta_ppo_actor_loss / ta_ppo_critic_loss   softmax, log-prob, entropy, ratio, clip, mean -- and their gradient
    ta_tinet_grad              all weight gradients -> the flat fp32 gradient buffer (the all-reduce operand)
    ta_adam_shadow             Adam on the flat buffers + the bf16 copy the next step reads
    ta_tinet_prep              per-step weight forms (folded conv1, permuted fc0, padded positionnet / head)

Parameters stay nn.Parameters with the reference's names (checkpoints interchange): their storage becomes views into one
flat fp32 buffer per network.  A step is ~90 launches for both networks instead of ~330 and is captured into a CUDA
graph by PPO.update exactly like the autograd step it replaces (ppo.py keeps that path: CPU, fp32, predictor nets).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from . import _capi
from . import conv1 as _c1


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _PrepArgs(C.Structure):
    _fields_ = [("w1", C.c_void_p), ("s_o", C.c_int64), ("s_c", C.c_int64), ("s_y", C.c_int64), ("s_x", C.c_int64),
                ("b1", C.c_void_p), ("w4", C.c_void_p), ("b4", C.c_void_p), ("fc0", C.c_void_p), ("fc0p", C.c_void_p),
                ("pos", C.c_void_p), ("pos16", C.c_void_p), ("head", C.c_void_p), ("head_b", C.c_void_p),
                ("head8", C.c_void_p), ("head_b8", C.c_void_p), ("nh", C.c_int)]


class _GradArgs(C.Structure):
    _fields_ = [("dw4", C.c_void_p), ("db4", C.c_void_p), ("g_w1", C.c_void_p), ("s_o", C.c_int64), ("s_c", C.c_int64),
                ("s_y", C.c_int64), ("s_x", C.c_int64), ("g_b1", C.c_void_p), ("src", C.c_void_p * 4), ("dst", C.c_void_p * 4),
                ("n", C.c_int64 * 4), ("fc0p", C.c_void_p), ("g_fc0", C.c_void_p), ("pos16", C.c_void_p), ("g_pos", C.c_void_p),
                ("head8", C.c_void_p), ("g_head", C.c_void_p), ("nh", C.c_int)]


class FusedNet:
    """One network (Net_PPO_actor or Net_PPO_critic) under the hand-scheduled step."""

    # parameter order of net.parameters(): cnn_base.{0,2,4,6}.{weight,bias}, positionnet, fc0, fc1, head
    NAMES = ("w1", "b1", "w2", "b2", "w3", "b3", "w4c", "b4c", "wpos", "bpos", "wfc0", "bfc0", "wfc1", "bfc1", "wh", "bh")
    IN_CH = 4    # input channels of the first convolution (FusedNet8: 8)

    def __init__(self, net: torch.nn.Module, kind: str, lr: float, eps: float):
        assert kind in ("actor", "critic")
        self.net, self.kind = net, kind
        self.lr, self.eps, self.betas = float(lr), float(eps), (0.9, 0.999)
        params = list(net.parameters())
        assert len(params) == 16 and params[0].shape == (64, self.IN_CH, 4, 4) and params[10].shape == (256, 2304)
        dev = params[0].device
        self.device = dev
        self.nh = params[14].shape[0]                       # 5 (actor) / 1 (critic)
        self.off, total = [], 0
        for p in params:
            self.off.append(total)
            total += (p.numel() + 63) // 64 * 64            # every tensor starts on a 256-byte boundary
        self.total = total
        self.P32 = torch.zeros(total, dtype=torch.float32, device=dev)
        self.G32 = torch.zeros(total, dtype=torch.float32, device=dev)
        self.M = torch.zeros(total, dtype=torch.float32, device=dev)
        self.V = torch.zeros(total, dtype=torch.float32, device=dev)
        self.P16 = torch.zeros(total, dtype=torch.bfloat16, device=dev)
        self.step_t = torch.zeros(1, dtype=torch.float32, device=dev)   # optimiser steps taken (float: what Adam's bias correction reads)
        self.params = params
        for p, off in zip(params, self.off):
            view = self._view(self.P32, off, p)
            view.copy_(p.data)
            p.data = view
            p.grad = self._view(self.G32, off, p)
        self.p16 = {n: self._view(self.P16, off, p) for n, p, off in zip(self.NAMES, params, self.off)}
        self.g32 = {n: self._view(self.G32, off, p) for n, p, off in zip(self.NAMES, params, self.off)}
        self.p32 = {n: p for n, p in zip(self.NAMES, params)}
        # derived per-step weight forms
        self.w4 = torch.empty((256, 16), dtype=torch.float32, device=dev)
        self.b4 = torch.empty((256,), dtype=torch.float32, device=dev)
        self.fc0p = torch.empty((256, 2304), dtype=torch.bfloat16, device=dev)
        self.pos16 = torch.empty((128, 16), dtype=torch.bfloat16, device=dev)
        self.head8 = torch.empty((8, 512), dtype=torch.bfloat16, device=dev)
        self.head_b8 = torch.empty((8,), dtype=torch.bfloat16, device=dev)
        # gradient staging in the forms the GEMMs produce
        self.g_fc0p = torch.empty((256, 2304), dtype=torch.bfloat16, device=dev)
        self.g_pos16 = torch.empty((128, 16), dtype=torch.bfloat16, device=dev)
        self.g_head8 = torch.empty((8, 512), dtype=torch.bfloat16, device=dev)
        self.g_w4c = torch.empty((256, 1152), dtype=torch.bfloat16, device=dev)
        self.g_wfc1 = torch.empty((512, 384), dtype=torch.bfloat16, device=dev)
        self.dw4 = torch.empty((256, 16), dtype=torch.float32, device=dev)
        self.db4 = torch.empty((256,), dtype=torch.float32, device=dev)
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.wimg2 = torch.empty(9 * 64 * 64, dtype=torch.bfloat16, device=dev)    # conv2 taps as the dgrad kernel's B operands
        import os
        self.tc_dgrad = os.environ.get("TA_CONV2_DGRAD_TC", "1") == "1"            # 0: cuDNN's merged-plane convolution
        self.stem_bwd_fused = os.environ.get("TA_STEM_BWD_FUSED", "1") == "1"      # 0: data gradient and conv1's weight gradient as two kernels
        self._scratch = {}
        self._L = _capi.lib()
        self.refresh()

    @staticmethod
    def _view(flat, off, p):
        n = p.numel()
        if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
            return flat[off:off + n].view(p.shape[0], p.shape[2], p.shape[3], p.shape[1]).permute(0, 3, 1, 2)
        return flat[off:off + n].view(p.shape)

    def _st(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _scr(self, key, rows, Cc):
        """Zero-initialised scratch of ta_relu_bwd_bias, one per call site (the two networks run on two streams)."""
        k = (key, rows, Cc)
        if k not in self._scratch:
            n = int(self._L.ta_relu_bwd_bias_scratch_floats(rows, Cc))
            self._scratch[k] = torch.zeros(n, dtype=torch.float32, device=self.device)
        return self._scratch[k]

    # ------------------------------------------------------------------ weights
    def refresh(self):
        """bf16 shadow and derived weight forms from the fp32 master (after construction, load_state_dict, broadcast)."""
        self.P16.copy_(self.P32)
        self._prep()

    def _prep(self):
        L, st = self._L, self._st()
        w1 = self.p32["w1"]
        a = _PrepArgs()
        a.w1, a.b1 = w1.data_ptr(), self.p32["b1"].data_ptr()
        a.s_o, a.s_c, a.s_y, a.s_x = w1.stride()
        a.w4, a.b4 = self.w4.data_ptr(), self.b4.data_ptr()
        a.fc0, a.fc0p = self.p16["wfc0"].data_ptr(), self.fc0p.data_ptr()
        a.pos, a.pos16 = self.p16["wpos"].data_ptr(), self.pos16.data_ptr()
        a.head, a.head_b = self.p16["wh"].data_ptr(), self.p16["bh"].data_ptr()
        a.head8, a.head_b8 = self.head8.data_ptr(), self.head_b8.data_ptr()
        a.nh = self.nh
        _capi.check(L.ta_tinet_prep(C.byref(a), st), "ta_tinet_prep")
        # conv2's data gradient runs on the tcgen05 per-class kernel (its operand image); conv3's as cuDNN's merged-plane convolution
        w2 = self.p16["w2"]
        _capi.check(L.ta_conv2_dgrad_prep(_ptr(w2), w2.stride(0), w2.stride(1), w2.stride(2), w2.stride(3), _ptr(self.wimg2), st),
                    "ta_conv2_dgrad_prep")
        if not self.tc_dgrad:
            self.pcw2 = _c1.parity_class_weights(self.p16["w2"])
        self.pcw3 = _c1.parity_class_weights(self.p16["w3"])

    # ------------------------------------------------------------------ one step
    def _relu_bwd(self, key, dy, ld, y2d, rows, Cc, db_name):
        """(dz, bias gradient written into the flat gradient buffer) of a ReLU layer; dy may be a column slice (ld)."""
        dz = torch.empty((rows, Cc), dtype=torch.bfloat16, device=self.device) if y2d is not None else None
        _capi.check(self._L.ta_relu_bwd_bias(_ptr(dy), ld, _ptr(y2d), _ptr(dz), rows, Cc, _ptr(self.g32[db_name]),
                                             _ptr(self._scr(key, rows, Cc)), self._st()), "ta_relu_bwd_bias")
        return dz

    def _finalise(self, conv1=False, dense=(), fc0=False, pos=False, head=False):
        """ta_tinet_grad on a subset of the gradient groups."""
        g32 = self.g32
        a = _GradArgs()
        if conv1:
            a.dw4, a.db4 = self.dw4.data_ptr(), self.db4.data_ptr()
            a.g_w1, a.g_b1 = g32["w1"].data_ptr(), g32["b1"].data_ptr()
            a.s_o, a.s_c, a.s_y, a.s_x = g32["w1"].stride()
        for k, (src, name) in enumerate(dense):
            a.src[k], a.dst[k], a.n[k] = src.data_ptr(), g32[name].data_ptr(), g32[name].numel()
        if fc0:
            a.fc0p, a.g_fc0 = self.g_fc0p.data_ptr(), g32["wfc0"].data_ptr()
        if pos:
            a.pos16, a.g_pos = self.g_pos16.data_ptr(), g32["wpos"].data_ptr()
        if head:
            a.head8, a.g_head = self.g_head8.data_ptr(), g32["wh"].data_ptr()
        a.nh = self.nh
        _capi.check(self._L.ta_tinet_grad(C.byref(a), self._st()), "ta_tinet_grad")

    def forward_backward(self, sb, pg16, loss_fn, reduce_fn=None, extra=None):
        """sb uint8 [B,4,289] codes, pg16 bf16 [B,16]; loss_fn(head_out bf16 [B,8], d_out bf16 [B,8], db_head fp32 view)
        launches the loss kernel.  Leaves every gradient in self.G32 and the mean loss in self.loss.
        reduce_fn(flat_slice, last) is the gradient all-reduce hook: it is called with the late layers' slice of G32
        (conv4 ... head: 91 % of the parameters) as soon as those gradients are final -- before the convolution stem's
        backward, which it then overlaps -- and with the stem's slice (last=True) at the end."""
        L, st, dev = self._L, self._st(), self.device
        B = sb.shape[0]
        p16, g32 = self.p16, self.g32
        bf = torch.bfloat16
        # ---------------- forward
        y1, mask = self._conv1_forward(sb, extra, B)
        y1v = y1.permute(0, 3, 1, 2)
        y2 = torch.cudnn_convolution_relu(y1v, p16["w2"], p16["b2"], [2, 2], [0, 0], [1, 1], 1)       # [B,64,16,16] channels-last
        y3 = torch.cudnn_convolution_relu(y2, p16["w3"], p16["b3"], [2, 2], [0, 0], [1, 1], 1)        # [B,128,7,7]
        y3n = y3.permute(0, 2, 3, 1)
        assert y3n.is_contiguous()
        cols4 = torch.empty((B * 9, 1152), dtype=bf, device=dev)
        _capi.check(L.ta_im2col_s2(_ptr(y3n), _ptr(cols4), B, 7, 7, 128, 3, st), "ta_im2col_s2")
        w4c = self.P16[self.off[6]:self.off[6] + 256 * 1152].view(256, 1152)                          # columns (ky, kx, c): channels-last memory
        y4 = torch._addmm_activation(p16["b4c"], cols4, w4c.t(), use_gelu=False)                      # [B*9,256]
        x5 = torch._addmm_activation(p16["bfc0"], y4.view(B, 2304), self.fc0p.t(), use_gelu=False)   # [B,256]
        pgz = torch._addmm_activation(p16["bpos"], pg16, self.pos16.t(), use_gelu=False)              # [B,128]
        x6 = torch.cat([x5, pgz], 1)                                                                  # [B,384]
        x7 = torch._addmm_activation(p16["bfc1"], x6, p16["wfc1"].t(), use_gelu=False)                # [B,512]
        out = torch.addmm(self.head_b8, x7, self.head8.t())                                           # [B,8]
        d_out = torch.empty((B, 8), dtype=bf, device=dev)
        loss_fn(out, d_out, g32["bh"])
        # ---------------- backward
        torch.mm(d_out.t(), x7, out=self.g_head8)
        dx7 = d_out @ self.head8                                                                      # [B,512]
        dz7 = self._relu_bwd("fc1", dx7, 512, x7, B, 512, "bfc1")
        torch.mm(dz7.t(), x6, out=self.g_wfc1)
        dx6 = dz7 @ p16["wfc1"]                                                                       # [B,384]
        dz5 = self._relu_bwd("fc0", dx6, 384, x5, B, 256, "bfc0")
        dzp = self._relu_bwd("pos", dx6[:, 256:], 384, pgz, B, 128, "bpos")
        torch.mm(dzp.t(), pg16, out=self.g_pos16)
        torch.mm(dz5.t(), y4.view(B, 2304), out=self.g_fc0p)
        dy4 = dz5 @ self.fc0p                                                                         # [B,2304] = [B*9,256]
        dz4 = self._relu_bwd("c4", dy4, 256, y4, B * 9, 256, "b4c")
        torch.mm(dz4.t(), cols4, out=self.g_w4c)
        # conv4 ... head: every gradient of the late layers is final -> flat buffer, all-reduce may start
        self._finalise(dense=((self.g_w4c, "w4c"), (self.g_wfc1, "wfc1")), fc0=True, pos=True, head=True)
        if reduce_fn is not None:
            reduce_fn(self.G32[self.off[6]:], False)
        dcols = dz4 @ w4c                                                                             # [B*9,1152]
        dy3 = torch.empty((B, 7, 7, 128), dtype=bf, device=dev)
        _capi.check(L.ta_col2im_s2(_ptr(dcols), _ptr(dy3), B, 7, 7, 128, 3, st), "ta_col2im_s2")
        dz3 = self._relu_bwd("c3", dy3, 128, y3n, B * 49, 128, "b3")
        dz3v = dz3.view(B, 7, 7, 128).permute(0, 3, 1, 2)
        gw3 = torch.ops.aten.convolution_backward(dz3v, y2, p16["w3"], None, [2, 2], [0, 0], [1, 1], False, [0, 0], 1,
                                                  [False, True, False])[1].contiguous(memory_format=torch.channels_last)
        p3 = F.conv2d(dz3v, self.pcw3, padding=1).permute(0, 2, 3, 1).contiguous()                    # merged parity planes
        dz2 = torch.empty((B, 16, 16, 64), dtype=bf, device=dev)
        y2n = y2.permute(0, 2, 3, 1)
        assert y2n.is_contiguous()
        # planes -> dense interleave + conv2's ReLU mask + conv2's bias gradient in one pass
        _capi.check(L.ta_planes_relu_bwd_bias(_ptr(p3), _ptr(y2n), _ptr(dz2), B, 16, 16, 64, _ptr(g32["b2"]),
                                              _ptr(self._scr("c2", B * 256, 64)), st), "ta_planes_relu_bwd_bias")
        dz2v = dz2.permute(0, 3, 1, 2)
        gw2 = torch.ops.aten.convolution_backward(dz2v, y1v, p16["w2"], None, [2, 2], [0, 0], [1, 1], False, [0, 0], 1,
                                                  [False, True, False])[1].contiguous(memory_format=torch.channels_last)
        self._stem_backward(dz2, dz2v, sb, extra, y1, mask, gw2, gw3, B)
        if reduce_fn is not None:
            reduce_fn(self.G32[:self.off[6]], True)
        return self.loss

    def _conv1_forward(self, sb, extra, B):
        """LUT decode + Upsample(4) + conv1 + bias + ReLU on tcgen05 with the ReLU bit mask: (y1 bf16 [B,33,33,64], mask)."""
        L, st, dev = self._L, self._st(), self.device
        y1 = torch.empty((B, 33, 33, 64), dtype=torch.bfloat16, device=dev)
        mask = torch.empty((B * 289 * 8,), dtype=torch.int32, device=dev)
        _capi.check(L.ta_conv1_fwd_mask(_ptr(sb), 1, sb.stride(0), _ptr(self.w4), _ptr(self.b4), B, _ptr(y1), _ptr(mask), st), "ta_conv1_fwd_mask")
        return y1, mask

    def _stem_backward(self, dz2, dz2v, sb, extra, y1, mask, gw2, gw3, B):
        """conv2's data gradient and conv1's weight / bias gradient; the stem's weight gradients into the flat fp32 buffer."""
        L, st, dev = self._L, self._st(), self.device
        bf = torch.bfloat16
        if self.tc_dgrad and self.stem_bwd_fused:
            # conv2's data gradient and conv1's weight / bias gradient in ONE tcgen05 kernel: the 605 MB of planes stay on the SM
            _capi.check(L.ta_conv2_dgrad_conv1_bwd(_ptr(dz2), _ptr(self.wimg2), _ptr(mask), _ptr(sb), 1, sb.stride(0), B,
                                                   _ptr(self.dw4), _ptr(self.db4), st), "ta_conv2_dgrad_conv1_bwd")
            self._finalise(conv1=True, dense=((gw2, "w2"), (gw3, "w3")))
            return
        if self.tc_dgrad:   # per-class tap lists on tcgen05 (warp-specialised), conv1's ReLU mask applied in the epilogue; class-major planes
            p2 = torch.empty((4, B * 289, 64), dtype=bf, device=dev)
            # (conv1's ReLU mask is applied by ta_conv1_bwd_planes, where it costs one multiply per word; in this kernel's
            # epilogue it quintuples the instructions per converted pair, and the epilogue warps are its critical path)
            _capi.check(L.ta_conv2_dgrad_planes(_ptr(dz2), _ptr(self.wimg2), None, B, 1, _ptr(p2), st), "ta_conv2_dgrad_planes")
        else:
            p2 = F.conv2d(dz2v, self.pcw2, padding=1).permute(0, 2, 3, 1).contiguous()
        _capi.check(L.ta_conv1_bwd_planes(_ptr(sb), 1, sb.stride(0), None, _ptr(mask), _ptr(p2), 1 if self.tc_dgrad else 0, B,
                                          _ptr(self.dw4), _ptr(self.db4), st), "ta_conv1_bwd_planes")
        # ---------------- the stem's weight gradients into the flat fp32 buffer
        self._finalise(conv1=True, dense=((gw2, "w2"), (gw3, "w3")))

    def adam(self, grad_scale: float = 1.0):
        b1, b2 = self.betas
        _capi.check(self._L.ta_adam_shadow(_ptr(self.P32), _ptr(self.G32), _ptr(self.M), _ptr(self.V), _ptr(self.P16), self.total,
                                           _ptr(self.step_t), self.lr, b1, b2, self.eps, float(grad_scale), self._st()), "ta_adam_shadow")
        self._prep()

    # ------------------------------------------------------------------ torch.optim.Adam-compatible state
    def export_adam_state(self, optimizer: torch.optim.Optimizer) -> dict:
        sd = optimizer.state_dict()
        state = {}
        if float(self.step_t.item()) > 0:
            for i, (p, off) in enumerate(zip(self.params, self.off)):
                state[i] = {"step": self.step_t.clone().reshape(()), "exp_avg": self._view(self.M, off, p).clone(),
                            "exp_avg_sq": self._view(self.V, off, p).clone()}
        return {"state": state, "param_groups": sd["param_groups"]}

    def import_adam_state(self, sd: dict):
        st = sd.get("state", {})
        if not st:
            return
        for i, (p, off) in enumerate(zip(self.params, self.off)):
            e = st.get(i, st.get(str(i)))
            if e is None:
                continue
            self._view(self.M, off, p).copy_(e["exp_avg"])
            self._view(self.V, off, p).copy_(e["exp_avg_sq"])
            self.step_t.fill_(float(e["step"]))


class FusedNet8(FusedNet):
    """FusedNet for the predictor agent's networks (Net_PPO_Predictor_actor / _critic, all_net.py:249-305): the first
    convolution has 8 input channels -- the 4 current frames (uint8 codes) and the 4 predicted ones (float).  The folded
    layer is linear in its input channels: forward = two passes of the 4-channel kernel joined by an addend
    (ta_conv1_fwd_add), backward = conv2's data gradient as class-major parity planes (ta_conv2_dgrad_planes) read by the
    tcgen05 weight-gradient kernel once per half with y1 as the ReLU mask (ta_conv1_bwd_planes); ta_tinet_prep /
    ta_tinet_grad fold / unfold each half of the weight through its element strides."""

    IN_CH = 8

    def __init__(self, net, kind, lr, eps):
        dev = next(net.parameters()).device
        self.w4b = torch.empty((256, 16), dtype=torch.float32, device=dev)
        self.b4_unused = torch.empty((256,), dtype=torch.float32, device=dev)
        self.b4_zero = torch.zeros((256,), dtype=torch.float32, device=dev)
        self.dw4b = torch.empty((256, 16), dtype=torch.float32, device=dev)
        self.db4b = torch.empty((256,), dtype=torch.float32, device=dev)
        self.gb1_unused = torch.empty((64,), dtype=torch.float32, device=dev)
        super().__init__(net, kind, lr, eps)

    def _prep(self):
        super()._prep()                       # folds input channels 0..3 (the strides step over all 8) and the other forms
        w1 = self.p32["w1"]
        a = _PrepArgs()
        a.w1, a.b1 = w1.data_ptr() + 4 * w1.stride(1) * 4, self.p32["b1"].data_ptr()      # channels 4..7
        a.s_o, a.s_c, a.s_y, a.s_x = w1.stride()
        a.w4, a.b4 = self.w4b.data_ptr(), self.b4_unused.data_ptr()
        a.fc0, a.fc0p = self.p16["wfc0"].data_ptr(), self.fc0p.data_ptr()                 # (rewritten with the same values)
        a.pos, a.pos16 = self.p16["wpos"].data_ptr(), self.pos16.data_ptr()
        a.head, a.head_b = self.p16["wh"].data_ptr(), self.p16["bh"].data_ptr()
        a.head8, a.head_b8 = self.head8.data_ptr(), self.head_b8.data_ptr()
        a.nh = self.nh
        _capi.check(self._L.ta_tinet_prep(C.byref(a), self._st()), "ta_tinet_prep")

    def _conv1_forward(self, sb, extra, B):
        L, st, dev = self._L, self._st(), self.device
        assert extra is not None and extra.dtype == torch.float32 and extra.shape == (B, 4, 289) and extra.is_contiguous()
        zb = torch.empty((B, 33, 33, 64), dtype=torch.bfloat16, device=dev)
        y1 = torch.empty_like(zb)
        _capi.check(L.ta_conv1_fwd_add(_ptr(extra), 0, extra.stride(0), _ptr(self.w4b), _ptr(self.b4_zero), B, None, 0, _ptr(zb), st),
                    "ta_conv1_fwd_add")
        _capi.check(L.ta_conv1_fwd_add(_ptr(sb), 1, sb.stride(0), _ptr(self.w4), _ptr(self.b4), B, _ptr(zb), 1, _ptr(y1), st),
                    "ta_conv1_fwd_add")
        return y1, None

    def _stem_backward(self, dz2, dz2v, sb, extra, y1, mask, gw2, gw3, B):
        L, st, dev = self._L, self._st(), self.device
        if self.tc_dgrad:
            p2 = torch.empty((4, B * 289, 64), dtype=torch.bfloat16, device=dev)
            _capi.check(L.ta_conv2_dgrad_planes(_ptr(dz2), _ptr(self.wimg2), None, B, 1, _ptr(p2), st), "ta_conv2_dgrad_planes")
        else:
            p2 = F.conv2d(dz2v, self.pcw2, padding=1).permute(0, 2, 3, 1).contiguous()
        cm = 1 if self.tc_dgrad else 0
        _capi.check(L.ta_conv1_bwd_planes(_ptr(sb), 1, sb.stride(0), _ptr(y1), None, _ptr(p2), cm, B, _ptr(self.dw4), _ptr(self.db4), st),
                    "ta_conv1_bwd_planes")
        _capi.check(L.ta_conv1_bwd_planes(_ptr(extra), 0, extra.stride(0), _ptr(y1), None, _ptr(p2), cm, B, _ptr(self.dw4b), _ptr(self.db4b), st),
                    "ta_conv1_bwd_planes")
        self._finalise(conv1=True, dense=((gw2, "w2"), (gw3, "w3")))     # channels 0..3 of w1, b1, w2, w3
        g_w1 = self.g32["w1"]
        a = _GradArgs()
        a.dw4, a.db4 = self.dw4b.data_ptr(), self.db4b.data_ptr()
        a.g_w1, a.g_b1 = g_w1.data_ptr() + 4 * g_w1.stride(1) * 4, self.gb1_unused.data_ptr()   # channels 4..7; the bias is counted once
        a.s_o, a.s_c, a.s_y, a.s_x = g_w1.stride()
        a.nh = self.nh
        _capi.check(L.ta_tinet_grad(C.byref(a), st), "ta_tinet_grad")


def supported(agent) -> bool:
    """The hand-scheduled step covers the plain PPO agent (4-frame TINet, FusedNet) and the predictor agent (8-channel
    first convolution, FusedNet8; TA_PPO_FUSED8=0 keeps it on the autograd path) on a GPU under bf16 autocast."""
    import os
    from . import ppo as _ppo
    if not (agent.device.type == "cuda" and agent.autocast and not agent.use_grad_clip):
        return False
    if type(agent) is _ppo.PPO:
        return isinstance(agent.actor, _ppo.Net_PPO_actor) and isinstance(agent.critic, _ppo.Net_PPO_critic)
    if type(agent).__name__ == "ppo_predictor" and os.environ.get("TA_PPO_FUSED8", "1") == "1":
        return bool(getattr(agent, "_pred_valid", False)) and agent.actor.bone1.cnn_base[0].weight.shape[1] == 8
    return False


def make(agent, net, kind, lr, eps):
    """FusedNet or FusedNet8 for one of the agent's networks."""
    cls = FusedNet8 if next(net.parameters()).shape[1] == 8 else FusedNet
    return cls(net, kind, lr, eps)
