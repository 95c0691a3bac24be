"""PPO + frozen frame predictor (BASELINE configs[4]) -- host-side mirror, vectorised.

Mirrors (reference file:line)
    soa/agent/net/all_net.py:7-137      Net_Encoder, LSTM, Net_Decoder
    soa/agent/net/all_net.py:249-305    Net_PPO_Predictor_actor / _critic (TINet with an 8-channel first conv)
    soa/agent/PPO_Predictor.py:25-194   ppo_predictor: pred_states, select_action, update
    soa/train_ppo_predictor.py:105-171  the 9-frame `pre_transition` records

Same class / parameter names and the same construction order (so reference checkpoints load and,
under the same torch seed, freshly built networks are bit-identical to the reference's;
tests/test_predictor_cpu.py checks both against fixtures produced by the reference).  Everything is
library code (cuDNN LSTM / conv, cuBLAS) on top of the device rollout of ppo.VecRollout: the env
step, featuriser and advantages are the CUDA library's.

PPO_Predictor.update only reads frames 0..4, a[:,0], r[:,0], a_logp[:,0] of its 9-frame records
(PPO_Predictor.py:124-163), and frames 0..4 of the record stored at step t are exactly the 5-frame
record the plain loop stores at step t-4 (train_ppo_predictor.py:123-171: the first four steps store
nothing, four padded records close the episode).  So the update runs on ppo.RolloutBuffer unchanged;
`pre_transition_records` materialises the full 9-frame records for the offline predictor training.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.distributions import Categorical

from .ppo import PPO, TINet, _weights_init, decode_matrix


class Net_Encoder(nn.Module):  # all_net.py:7-51
    def __init__(self):
        super().__init__()
        self.cnn_base = nn.Sequential(
            nn.Conv2d(1, 16, kernel_size=4, stride=2), nn.BatchNorm2d(16), nn.ReLU(),     # (16, 33, 33)
            nn.Conv2d(16, 16, kernel_size=5, stride=4), nn.BatchNorm2d(16), nn.ReLU(),    # (16, 8, 8)
            nn.Conv2d(16, 64, kernel_size=2, stride=2), nn.BatchNorm2d(64), nn.ReLU(),    # (64, 4, 4)
        )
        self.apply(_weights_init)
        self.upsamplingnearest = nn.UpsamplingNearest2d(scale_factor=4)
        self.device = None

    def forward(self, state_matrix):
        B, T, _ = state_matrix.shape
        x = state_matrix.reshape(-1, 1, 289).contiguous().view(-1, 1, 17, 17)
        up = self.upsamplingnearest(x).float()
        z = self.cnn_base(up)
        return z.view(-1, T, 64, 4, 4), up.view(-1, T, 1, 68, 68)


class LSTM(nn.Module):  # all_net.py:53-98
    """nn.LSTM(1024, 1024, 3) over the 4 encoded frames, then 3 steps fed with its own top-layer output.

    On the GPU under autocast the forward does not go through cuDNN's RNN (measured: 81 ms per 8192-env step, 36 TFLOP/s)
    but through `_fast_forward`: per layer ONE bf16 GEMM for the input gates of all four teacher-forced steps and one per
    step for the recurrent gates (fp32 out), the cell update by the library's `ta_lstm_gates` kernel, and for the
    extrapolation steps one K = 2048 GEMM over the concatenated [x, h] operand whose h half the gate kernel writes in
    place.  Same arithmetic as torch.nn.LSTM (gate order i, f, g, o; c kept in fp32); `TA_LSTM_FAST=0` selects cuDNN."""

    def __init__(self):
        super().__init__()
        self.extrap_t = 4
        self.nt = 8
        self.recurrent_model = nn.LSTM(1024, 1024, num_layers=3, batch_first=True)
        self.h_0 = torch.zeros(3, 1024)
        self.c_0 = torch.zeros(3, 1024)
        self.device = None
        self._fast_cache = None

    def forward(self, z_content):
        B, T, D, W, H = z_content.shape
        z_content = z_content.reshape(B, T, D * W * H)
        if (z_content.is_cuda and not torch.is_grad_enabled() and torch.is_autocast_enabled("cuda")
                and os.environ.get("TA_LSTM_FAST", "1") == "1"):
            z = self._fast_forward(z_content, torch.bfloat16)
            return z.reshape(B, self.nt - 1, D, W, H), z_content
        h_0 = self.h_0.unsqueeze(1).repeat(1, B, 1).to(z_content.device, z_content.dtype)
        c_0 = self.c_0.unsqueeze(1).repeat(1, B, 1).to(z_content.device, z_content.dtype)
        z_past, (h_n, c_n) = self.recurrent_model(z_content, (h_0, c_0))
        z_n = z_past[:, -1].unsqueeze(1)
        prediction = []
        for _ in range(self.nt - 4 - 1):
            z_n, (h_n, c_n) = self.recurrent_model(z_n, (h_n, c_n))
            prediction.append(z_n)
        z = torch.cat([z_past, torch.cat(prediction, 1)], 1)
        return z.reshape(B, self.nt - 1, D, W, H), z_content

    # ---- the hand-scheduled forward --------------------------------------------------------------------------------
    def _fast_weights(self, dtype):
        """Per layer (W_ih^T, W_hh^T, [W_ih | W_hh]^T as views of `dtype` copies, b_ih + b_hh in fp32); rebuilt when a
        parameter changed (load_state_dict copies in place and bumps the version counters)."""
        rm = self.recurrent_model
        ps = [getattr(rm, f"{n}_l{l}") for l in range(rm.num_layers) for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
        key = (dtype, tuple((p.data_ptr(), p._version) for p in ps))
        if self._fast_cache is None or self._fast_cache[0] != key:
            layers = []
            for l in range(rm.num_layers):
                w_ih, w_hh, b_ih, b_hh = (p.detach() for p in ps[4 * l:4 * l + 4])
                w_cat = torch.cat([w_ih, w_hh], 1).to(dtype).contiguous()         # [4H, in + H]
                layers.append((w_ih.to(dtype).contiguous().t(), w_hh.to(dtype).contiguous().t(), w_cat.t(),
                               (b_ih.float() + b_hh.float()).contiguous()))
            self._fast_cache = (key, layers)
        return self._fast_cache[1]

    @staticmethod
    def _mm32(a, b):
        """a @ b with an fp32 result (tensor cores with fp32 accumulate when the operands are bf16)."""
        if a.is_cuda and a.dtype != torch.float32:
            return torch.mm(a, b, out_dtype=torch.float32)
        return torch.mm(a.float(), b.float())

    @staticmethod
    def _gates(gx, gh, bias, c, h_out):
        """The cell update: c (fp32 [B,H]) in place, h into h_out ([B,H] view whose rows may be strided)."""
        B, H = c.shape
        if gx.is_cuda and h_out.dtype == torch.bfloat16:
            import ctypes as C
            from . import _capi
            assert gx.is_contiguous() and (gh is None or gh.is_contiguous()) and c.is_contiguous() and h_out.stride(1) == 1
            _capi.check(_capi.lib().ta_lstm_gates(
                C.c_void_p(gx.data_ptr()), None if gh is None else C.c_void_p(gh.data_ptr()), C.c_void_p(bias.data_ptr()),
                C.c_void_p(c.data_ptr()), C.c_void_p(h_out.data_ptr()), h_out.stride(0), B, H,
                C.c_void_p(torch.cuda.current_stream(gx.device).cuda_stream)), "ta_lstm_gates")
            return
        pre = gx + bias if gh is None else gx + gh + bias
        i, f, g, o = pre.split(H, 1)
        c.copy_(torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g))
        h_out.copy_((torch.sigmoid(o) * torch.tanh(c)).to(h_out.dtype))

    @torch.no_grad()
    def _fast_forward(self, z_content, dtype=torch.bfloat16):
        """z_content [B, T, 1024] -> [B, T + nt - 5, 1024] (the T outputs, then the nt - 5 self-fed steps), in `dtype`."""
        B, T, D = z_content.shape
        H = self.recurrent_model.hidden_size
        dev = z_content.device
        with torch.autocast(device_type=dev.type, enabled=False):
            layers = self._fast_weights(dtype)
            x_seq = z_content.transpose(0, 1).to(dtype).contiguous()              # [T, B, D], time-major
            n_ext = self.nt - 4 - 1
            xh, cs = [], []
            for l, (w_ih_t, w_hh_t, w_cat_t, bias) in enumerate(layers):
                gx = self._mm32(x_seq.view(T * B, -1), w_ih_t).view(T, B, 4 * H)  # input gates of all T steps: one GEMM
                c = self.c_0[l].to(dev, torch.float32).expand(B, H).contiguous()
                hl = self.h_0[l].to(dev, dtype).expand(B, H).contiguous()
                out = torch.empty((T, B, H), dtype=dtype, device=dev)
                for t in range(T):
                    self._gates(gx[t], self._mm32(hl, w_hh_t), bias, c, out[t])
                    hl = out[t]
                buf = torch.empty((B, x_seq.shape[2] + H), dtype=dtype, device=dev)   # the [x | h] operand of the self-fed steps
                buf[:, -H:] = hl
                xh.append(buf); cs.append(c)
                x_seq = out
            z = torch.empty((T + n_ext, B, H), dtype=dtype, device=dev)
            z[:T] = x_seq
            x = x_seq[T - 1]
            for s in range(n_ext):
                for l, (_, _, w_cat_t, bias) in enumerate(layers):
                    xh[l][:, :-H] = x
                    self._gates(self._mm32(xh[l], w_cat_t), None, bias, cs[l], xh[l][:, -H:])
                    x = xh[l][:, -H:]
                z[T + s] = x
            return z.transpose(0, 1).contiguous()


class Net_Decoder(nn.Module):  # all_net.py:100-137
    def __init__(self):
        super().__init__()
        self.cnn_base = nn.Sequential(
            nn.ConvTranspose2d(64, 16, kernel_size=2, stride=2), nn.ReLU(),
            nn.ConvTranspose2d(16, 16, kernel_size=5, stride=4), nn.ReLU(),
            nn.ConvTranspose2d(16, 1, kernel_size=4, stride=2),
        )
        self.apply(_weights_init)
        self.pool = nn.AvgPool2d(4, stride=4)

    def forward(self, state_matrix):
        B, T, D, W, H = state_matrix.shape
        full = self.cnn_base(state_matrix.contiguous().view(-1, D, W, H))
        pooled = self.pool(full).view(-1, 1, 289).reshape(-1, T, 289)
        return pooled, full.view(-1, T, 1, 68, 68)


class Net_PPO_Predictor_actor(nn.Module):  # all_net.py:249-276
    def __init__(self):
        super().__init__()
        self.bone1 = TINet()
        self.bone1.cnn_base[0] = nn.Conv2d(8, 64, kernel_size=4, stride=2)
        self.A = nn.Linear(512, 5)
        self.apply(_weights_init)

    def forward(self, state_matrix, position, goal):
        return torch.softmax(self.A(self.bone1(state_matrix, position, goal)).float(), dim=1)


class Net_PPO_Predictor_critic(nn.Module):  # all_net.py:278-305
    def __init__(self):
        super().__init__()
        self.bone2 = TINet()
        self.bone2.cnn_base[0] = nn.Conv2d(8, 64, kernel_size=4, stride=2)
        self.V = nn.Linear(512, 1)
        self.apply(_weights_init)

    def forward(self, state_matrix, position, goal):
        return self.V(self.bone2(state_matrix, position, goal)).float()


class ppo_predictor(PPO):
    """PPO_Predictor.py:25-194: PPO whose actor / critic see the 4 current frames plus 4 frames
    predicted by a frozen Encoder -> LSTM -> Decoder.  Only the actor and critic are trained
    (the encoder / decoder / predictor optimisers exist in the reference but never step)."""

    def __init__(self, device="cpu", autocast: Optional[bool] = None, flat_grads: bool = True):
        self.device = torch.device(device)
        # construction order of PPO_Predictor.py:32-36 (it fixes the RNG stream of the initial weights)
        actor = Net_PPO_Predictor_actor()
        critic = Net_PPO_Predictor_critic()
        self.encoder = Net_Encoder().to(self.device)
        self.decoder = Net_Decoder().to(self.device)
        self.predictor = LSTM().to(self.device)
        super().__init__(device=device, autocast=autocast, flat_grads=flat_grads, _nets=(actor, critic))
        self.encoder.device = self.predictor.device = self.device
        # h_0 / c_0 are plain attributes in the reference (moved on every forward); keep them on the device so the
        # forward has no host-to-device copy (not capturable in the CUDA graph of the optimiser step)
        self.predictor.h_0, self.predictor.c_0 = self.predictor.h_0.to(self.device), self.predictor.c_0.to(self.device)
        self.encoder_lr = self.decoder_lr = self.predictor_lr = 0.00001
        self._pred_table, self._pred_valid = None, False
        self._stack_cache = None

    def load_predictor(self, state):
        """train_ppo_predictor.py:77-81: the pre-trained predictor stack from a checkpoint dict."""
        self.encoder.load_state_dict(state["model_encoder"])
        self.decoder.load_state_dict(state["model_decoder"])
        self.predictor.load_state_dict(state["model_predictor"])

    # ---- the fused inference path (csrc/ta_pred.cuh) -------------------------------------------------------------------
    def _stack_arrays(self):
        """The arrays ta_pred_encoder / ta_pred_decoder read, from the modules' parameters (fp32, on the device): tap-major
        convolution weights, eval-mode BatchNorm + bias folded into scale / shift, and the last transposed convolution
        folded with the 4x4 average pool into a 3x3 stride-2 padding-1 convolution.  Rebuilt when a parameter or a
        running statistic changed."""
        enc, dec = self.encoder.cnn_base, self.decoder.cnn_base
        src = [t for m in (self.encoder, self.decoder) for t in list(m.parameters()) + list(m.buffers())]
        key = tuple((t.data_ptr(), t._version) for t in src)
        if self._stack_cache is not None and self._stack_cache[0] == key:
            return self._stack_cache[1]

        def bn_fold(conv, bn):
            s = bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)
            return s.contiguous(), ((conv.bias.float() - bn.running_mean.float()) * s + bn.bias.float()).contiguous()

        with torch.no_grad():
            e = {"w1": enc[0].weight.float().reshape(16, 16).contiguous(),
                 "w2": enc[3].weight.float().permute(2, 3, 1, 0).contiguous(),      # [co][ci][ky][kx] -> [ky][kx][ci][co]
                 "w3": enc[6].weight.float().permute(2, 3, 1, 0).contiguous()}
            (e["s1"], e["t1"]), (e["s2"], e["t2"]), (e["s3"], e["t3"]) = bn_fold(enc[0], enc[1]), bn_fold(enc[3], enc[4]), bn_fold(enc[6], enc[7])
            d = {"w1": dec[0].weight.float().permute(2, 3, 0, 1).contiguous(),      # [ci][co][ky][kx] -> [ky][kx][ci][co]
                 "b1": dec[0].bias.float().contiguous(),
                 "w2": dec[2].weight.float().permute(2, 3, 0, 1).contiguous(),
                 "b2": dec[2].bias.float().contiguous()}
            d["w3"], b3 = fold_decoder_tail(dec[4].weight.float(), dec[4].bias.float())
            d["b3"] = float(b3)       # (one host read when the cache is built, none per call)
        self._stack_cache = (key, (e, d))
        return e, d

    def _pred_states_fused(self, state_matrix):
        """[B,4,289] uint8 codes or float LUT values -> [B,4,289] float32: encoder kernel, the LSTM's hand-scheduled
        forward, decoder kernel."""
        import ctypes as C
        from . import _capi
        L = _capi.lib()
        dev = state_matrix.device
        B = state_matrix.shape[0]
        x = state_matrix.contiguous() if state_matrix.dtype == torch.uint8 else state_matrix.float().contiguous()
        e, d = self._stack_arrays()
        ptr = lambda t: C.c_void_p(t.data_ptr())
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        z_c = torch.empty((B, 4, 1024), dtype=torch.bfloat16, device=dev)
        _capi.check(L.ta_pred_encoder(ptr(x), 1 if x.dtype == torch.uint8 else 0, B * 4, ptr(e["w1"]), ptr(e["s1"]), ptr(e["t1"]),
                                      ptr(e["w2"]), ptr(e["s2"]), ptr(e["t2"]), ptr(e["w3"]), ptr(e["s3"]), ptr(e["t3"]), ptr(z_c), st),
                    "ta_pred_encoder")
        z_pred = self.predictor._fast_forward(z_c, torch.bfloat16)                 # [B, 7, 1024]
        z_in = z_pred[:, 3:7].contiguous()
        out = torch.empty((B, 4, 289), dtype=torch.float32, device=dev)
        _capi.check(L.ta_pred_decoder(ptr(z_in), B * 4, ptr(d["w1"]), ptr(d["b1"]), ptr(d["w2"]), ptr(d["b2"]), ptr(d["w3"]),
                                      float(d["b3"]), ptr(out), st), "ta_pred_decoder")
        return out

    @torch.no_grad()
    def pred_states(self, state_matrix):
        """PPO_Predictor.py:70-83: [B,4,289] current frames -> [B,4,289] predicted next frames (+ the upsampled input and
        the full-resolution decoder output, which no caller of the PPO loop uses: None on the fused GPU path)."""
        if state_matrix.is_cuda and self.autocast and os.environ.get("TA_PRED_FUSED", "1") == "1":
            return self._pred_states_fused(state_matrix), None, None
        if state_matrix.dtype == torch.uint8:
            state_matrix = decode_matrix(state_matrix)
        states_pre = state_matrix.reshape(-1, 1, 289)
        self.encoder.eval(); self.decoder.eval(); self.predictor.eval()
        with self._amp():
            z_c, z_c_upsample = self.encoder(states_pre)
            z_c = z_c.view(-1, 4, 64, 4, 4)
            z_pred, _ = self.predictor(z_c)
            states_head, states_head_pool = self.decoder(z_pred[:, 3:7])
        return states_head.float(), z_c_upsample, states_head_pool

    def _cat(self, frames):
        if frames.dtype == torch.uint8:
            frames = decode_matrix(frames)
        frames = frames.float()
        return torch.cat([frames, self.pred_states(frames)[0].detach()], 1)   # [B,8,289]

    def _net_in(self, frames, rows=None, lo=None):
        """PPO_Predictor.py:100-103 / :133-139 / :149-150: current frames + the predicted ones.  Inside update() the
        predictions of a record's frames 0..3 come from the per-buffer table of _begin_update."""
        if lo == 0 and rows is not None and self._pred_valid:
            if frames.dtype == torch.uint8:
                frames = decode_matrix(frames)
            return torch.cat([frames.float(), self._pred_table[rows].float()], 1)
        return self._cat(frames)

    def _begin_update(self, s):
        """The predictor stack is frozen (its optimisers never step, PPO_Predictor.py:124-163), so the predicted frames of a
        record do not change during update(): they are computed ONCE per buffer row here (bf16 [B,4,289], in place in a
        table that persists across updates) instead of once per critic pass and per minibatch of every epoch --
        1 + K_epochs fewer passes of Encoder -> LSTM -> Decoder over the buffer.  TA_PRED_CACHE=0 recomputes them per use."""
        self._pred_valid = False
        B = s.shape[0]
        if not s.is_cuda or os.environ.get("TA_PRED_CACHE", "1") != "1" or B * 4 * 289 * 2 > 32e9:
            return
        if self._pred_table is None or self._pred_table.shape[0] != B or self._pred_table.device != s.device:
            self._pred_table = torch.empty((B, 4, 289), dtype=torch.bfloat16, device=s.device)
        chunk = 16384
        for i in range(0, B, chunk):
            self._pred_table[i:i + chunk] = self.pred_states(s[i:i + chunk, 0:4])[0]
        self._pred_valid = True

    def _end_update(self):
        self._pred_valid = False

    def state_dict(self, i_ep: int = 0):
        d = super().state_dict(i_ep)
        d.update(model_encoder=self.encoder.state_dict(), model_decoder=self.decoder.state_dict(),
                 model_predictor=self.predictor.state_dict())
        return d


def fold_decoder_tail(w, b):
    """ConvTranspose2d(16, 1, 4, 2) followed by AvgPool2d(4, 4) (all_net.py:106-108,114) as ONE convolution of the 33 x 33 map:
    pooled[Y][X] averages the outputs y' = 4Y..4Y+3, and y' = 2i + ky, so input row i = 2Y - 1 contributes through ky in {2, 3},
    i = 2Y through ky in {0..3}, i = 2Y + 1 through ky in {0, 1}: a 3 x 3 kernel, stride 2, padding 1 whose taps are sums of the
    4 x 4 kernel's divided by 16.  w [16][1][4][4], b [1] -> (taps [3][3][16] contiguous, bias)."""
    sets = ((2, 3), (0, 1, 2, 3), (0, 1))
    w = w.reshape(16, 4, 4)
    taps = torch.stack([torch.stack([sum(w[:, ky, kx] for ky in sets[dy] for kx in sets[dx]) / 16.0 for dx in range(3)])
                        for dy in range(3)])                                          # [3][3][16]
    return taps.contiguous(), b.reshape(())


def pre_transition_records(ended: torch.Tensor):
    """The 9-frame `pre_transition` records of train_ppo_predictor.py:105-171 as an index table over a
    [T,N] rollout.  ended bool/uint8 [T,N] (terminated | truncated).  Episodes are the segments between
    `ended` marks.  Returns (index int64 [M,9], env int64 [M]): record m consists of the frames produced
    by steps index[m, 0..8] of env[m], where -1 stands for the frame MiniGridEnv.reset leaves behind
    (predata_reset tiles it nine times, env_buffer.py:430-437) -- also used for the unknown history of
    an episode that began before the window.  For every step t >= 4 of an episode the window t-8..t
    (`if t > 3`, :140-142), plus four closing records that repeat the terminal frame (:145-160)."""
    T, N = ended.shape
    ended = ended.bool()
    dev = ended.device
    t_idx = torch.arange(T, device=dev).view(T, 1).expand(T, N)
    prev_end = torch.where(ended, t_idx + 1, torch.zeros_like(t_idx))
    start = torch.cummax(torch.cat([torch.zeros(1, N, dtype=torch.long, device=dev), prev_end[:-1]]), dim=0).values
    age = t_idx - start                                         # 0-based step inside the episode
    offs = torch.arange(-8, 1, device=dev).view(1, 1, 9)
    e_idx = torch.arange(N, device=dev).view(1, N).expand(T, N)
    recs, envs = [], []

    def clip(w):
        return torch.where(w < start.unsqueeze(-1), torch.full_like(w, -1), w)

    m = age >= 4
    recs.append(clip(t_idx.unsqueeze(-1) + offs)[m]); envs.append(e_idx[m])
    for k in range(1, 5):                                       # the 4 closing records
        w = torch.minimum(t_idx.unsqueeze(-1) + offs + k, t_idx.unsqueeze(-1))
        recs.append(clip(w)[ended]); envs.append(e_idx[ended])
    return torch.cat(recs), torch.cat(envs)


def gather_records(frames: torch.Tensor, reset_frame: torch.Tensor, index: torch.Tensor, env: torch.Tensor):
    """frames [T,N,...] + the reset frame [...] -> records [M,9,...] for an index table of pre_transition_records."""
    ext = torch.cat([frames, reset_frame.expand(1, *frames.shape[1:]).to(frames.dtype)], 0)   # row T = reset frame
    t = torch.where(index < 0, torch.full_like(index, frames.shape[0]), index)
    return ext[t, env.unsqueeze(-1)]
