"""Vectorised PPO for the batched Twoarmy env -- host-side mirror of the reference's agent.

Mirrors (reference file:line, paths relative to the reference root)
    soa/agent/net/all_net.py:139-247   TINet, Net_PPO_actor, Net_PPO_critic
    soa/agent/PPO.py:41-161            PPO: hyper-parameters, select_action, update
    soa/train_ppo.py:93-160            rollout buffer record, episode loop

Same class / attribute / parameter names (`actor.bone1.cnn_base.0.weight`, ...), same
initialisation order and the same arithmetic, so a reference checkpoint loads with
`load_state_dict` and, under the same torch seed, the networks are bit-identical to the
reference's (tests/test_ppo_cpu.py checks both against fixtures produced by the reference).
What changes is the batching: select_action takes N envs at once, update can use large
minibatches, bf16 autocast + channels_last on the GPU (cuDNN / cuBLAS tensor-core kernels: per
BASELINE.json's north_star the policy/value GEMMs and convolutions are the only tensor-core
work on this path and stay library calls), and gradients are all-reduced over
torch.distributed (NCCL on GPUs, gloo in the CPU tests) when a process group is up.

The env step, observation, featuriser and advantage arithmetic all run in the CUDA library
(include/twoarmy_b200.h); nothing here re-implements them.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.distributions import Categorical
from torch.utils.data.sampler import BatchSampler, SubsetRandomSampler

# Env_transact.matrix_env (soa/env_buffer.py:300-318) as a table over the compact codes the
# featuriser kernel emits: 0 empty/goal 0.9, 1 wall -0.9, 2 ball -0.5, 4 agent 0.3
MATRIX_LUT = (0.9, -0.9, -0.5, 0.0, 0.3)
# Env_transact.env_action (soa/env_buffer.py:364-376): policy index -> env action
POLICY_TO_ENV_ACTION = (0, 1, 2, 3, 6)


def _weights_init(m):  # all_net.py:162-172 (identical in the three classes)
    if isinstance(m, nn.Linear):
        nn.init.xavier_normal_(m.weight)
        nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.Conv2d):
        nn.init.xavier_uniform_(m.weight, gain=nn.init.calculate_gain("relu"))
        nn.init.constant_(m.bias, 0.1)
    elif isinstance(m, nn.BatchNorm2d):
        nn.init.constant_(m.weight, 1)
        nn.init.constant_(m.bias, 0)


class TINet(nn.Module):
    """all_net.py:139-189: 4 frames of 17x17 -> nearest x4 -> 4 convs -> fc0; positions+goal ->
    positionnet; concat -> fc1 -> 512 features."""

    def __init__(self, in_frames: int = 4):
        super().__init__()
        self.cnn_base = nn.Sequential(
            nn.Conv2d(in_frames, 64, kernel_size=4, stride=2), nn.ReLU(),   # (64, 33, 33)
            nn.Conv2d(64, 64, kernel_size=3, stride=2), nn.ReLU(),          # (64, 16, 16)
            nn.Conv2d(64, 128, kernel_size=4, stride=2), nn.ReLU(),         # (128, 7, 7)
            nn.Conv2d(128, 256, kernel_size=3, stride=2), nn.ReLU(),        # (256, 3, 3)
            nn.Flatten(),
        )
        self.positionnet = nn.Linear(10, 128)
        self.fc0 = nn.Linear(2304, 256)
        self.fc1 = nn.Linear(256 + 128, 512)
        self.upsamplingnearest = nn.UpsamplingNearest2d(scale_factor=4)
        self.apply(_weights_init)

    # Sel[p][d][k]: which of the 4 kernel taps k of one axis land on input offset d for output phase p
    # (even outputs see one input pixel through all 4 taps, odd outputs two pixels through 2 taps each)
    _SEL = ((( 1., 1., 1., 1.), (0., 0., 0., 0.)), ((1., 1., 0., 0.), (0., 0., 1., 1.)))
    # Off by default: built from torch ops (cat / GEMM / interleaving copy) the folded layer measured
    # 22.0 ms per optimiser step against 19.9 ms for upsample + cuDNN (B = 4096, two nets); it only
    # pays once the interleave is fused into the GEMM epilogue.
    fold_conv1 = False
    # the hand-written kernel pair for this layer (csrc/ta_conv1.cuh), used on the GPU under bf16 autocast
    fused_conv1 = True
    # conv1 + conv2 + conv3 in one autograd node (the data gradients of conv2 / conv3 as four parity-class
    # convolutions each; conv2's feed the conv1 weight-gradient kernel directly)
    fused_stem = True
    # Linear + ReLU layers through conv1._LinearReLU on the GPU under bf16 autocast
    fused_linear = True
    # data gradients of conv2 / conv3 as GEMM + col2im (csrc/ta_conv1.cuh) instead of cuDNN's strided dgrad
    gemm_dgrad = True

    def _conv1_folded(self, x):
        """UpsamplingNearest2d(4) + Conv2d(4, 64, k=4, s=2) + ReLU (all_net.py:142-143,157,180-181) folded
        exactly: output pixel (2m+py, 2n+px) only sees input pixels
        (m+dy, n+dx), dy, dx in {0,1}, through sums of kernel taps.  So the layer is one GEMM over 2x2
        patches of the 17x17 frame (K = 16) producing the four phases of 64 channels, interleaved by
        a single copy -- no 68x68 tensor, no channel-padded cuDNN kernel.  x [B,4,17,17] ->
        [B,64,33,33] channels_last."""
        conv1 = self.cnn_base[0]
        B = x.shape[0]
        sel = torch.tensor(self._SEL, dtype=conv1.weight.dtype, device=x.device)           # [p, d, k]
        w = torch.einsum("ocyx,pdy,qex->pqodec", conv1.weight, sel, sel).reshape(256, 16)  # rows (py,px,o), cols (dy,dx,c)
        xn = F.pad(x.permute(0, 2, 3, 1), (0, 0, 0, 1, 0, 1))                              # [B,18,18,4]
        cols = torch.cat([xn[:, dy:dy + 17, dx:dx + 17, :] for dy in range(2) for dx in range(2)], dim=3)  # [B,17,17,16]
        if torch.is_autocast_enabled():
            cols, w = cols.to(torch.bfloat16), w.to(torch.bfloat16)
        y = F.linear(cols.reshape(B * 289, 16), w.to(cols.dtype), conv1.bias.repeat(4).to(cols.dtype))   # [B*289, 256]
        y = y.view(B, 17, 17, 2, 2, 64).permute(0, 1, 3, 2, 4, 5).reshape(B, 34, 34, 64)[:, :33, :33, :]
        return torch.relu(y.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last))

    _c4_idx = {}

    @classmethod
    def _conv4_index(cls, device):
        key = str(device)
        if key not in cls._c4_idx:
            idx = [(2 * oy + ky) * 7 + (2 * ox + kx) for oy in range(3) for ox in range(3) for ky in range(3) for kx in range(3)]
            cls._c4_idx[key] = torch.tensor(idx, dtype=torch.int64, device=device)
        return cls._c4_idx[key]

    def forward(self, state_matrix, position, goal):
        """state_matrix [B,4,289]: float matrix_env values, or (GPU) the uint8 featuriser codes."""
        B, T, _ = state_matrix.shape
        position = position.contiguous().view(-1, 8)
        fast = state_matrix.is_cuda and torch.is_autocast_enabled() and self.fused_linear
        if fast:  # bias + ReLU in the GEMM epilogue, bias gradient from the channel-sum kernel (conv1._LinearReLU)
            from . import conv1 as _c1
            position_goal = _c1.linear_relu(torch.cat([position, goal], 1).to(torch.bfloat16), self.positionnet)
        else:
            position_goal = torch.relu(self.positionnet(torch.cat([position, goal], 1)))
        if not state_matrix.is_cuda:  # the reference's layer sequence, verbatim
            if state_matrix.dtype == torch.uint8:
                state_matrix = decode_matrix(state_matrix)
            x = self.upsamplingnearest(state_matrix.contiguous().view(-1, T, 17, 17))
            x = torch.relu(self.fc0(self.cnn_base(x)))
        else:
            stem = False
            if self.fused_conv1 and torch.is_autocast_enabled() and T == 4:
                # LUT decode + upsample + conv1 + bias + ReLU in one kernel, straight from codes or floats
                from . import conv1 as _c1
                ok = state_matrix.stride(2) == 1 and state_matrix.stride(1) == 289
                xin = state_matrix if ok else state_matrix.contiguous()
                stem = self.fused_stem and self.gemm_dgrad
                # conv1 + conv2 + conv3 as one autograd node: the data gradients stay in parity planes (conv1._Stem)
                x = _c1.stem_relu(xin, self.cnn_base[0], self.cnn_base[2], self.cnn_base[4]) if stem else _c1.conv1_relu(xin, self.cnn_base[0])
            elif (self.fused_conv1 and self.fused_stem and self.gemm_dgrad and torch.is_autocast_enabled() and T == 8
                  and state_matrix.dtype == torch.float32 and os.environ.get("TA_STEM8", "1") == "1"):
                # 8 input channels (the predictor agent's networks): the folded layer as two 4-channel passes (conv1._Stem8)
                from . import conv1 as _c1
                stem = True
                x = _c1.stem8_relu(state_matrix.contiguous(), self.cnn_base[0], self.cnn_base[2], self.cnn_base[4])
            else:
                if state_matrix.dtype == torch.uint8:
                    state_matrix = decode_matrix(state_matrix)
                x = state_matrix.contiguous().view(-1, T, 17, 17)
                if self.fold_conv1:
                    x = self._conv1_folded(x)
                else:
                    x = self.upsamplingnearest(x).contiguous(memory_format=torch.channels_last)
                    x = self.cnn_base[:2](x)
            if self.gemm_dgrad and x.dtype == torch.bfloat16:
                from . import conv1 as _c1
                if not stem:
                    x = _c1.conv_s2_relu(x, self.cnn_base[2])
                    x = _c1.conv_s2_relu(x, self.cnn_base[4])
            else:
                x = self.cnn_base[2:6](x)
            # The last conv (128 -> 256, 3x3 stride 2 on 7x7) as an explicit im2col + cuBLAS GEMM: for
            # this shape cuDNN picks a kernel that takes 1.8 ms fwd+bwd at B = 4096, the GEMM 0.2 ms
            # (scripts/conv_gemm_probe.py).  Same arithmetic, same parameters.
            conv4 = self.cnn_base[6]
            # im2col as ONE gather (torch's unfold loops over the batch): row (o, k) of the index = input pixel
            # (2*oy+ky, 2*ox+kx), so the columns end up ordered (ky, kx, c); its backward is the col2im kernel
            xn = x.permute(0, 2, 3, 1)                                        # [B,7,7,128], a view (channels_last)
            if x.dtype == torch.bfloat16:
                from . import conv1 as _c1
                cols = _c1.im2col_s2(xn, self._conv4_index(x.device), 3)
            else:
                cols = xn.reshape(B, 49, 128).index_select(1, self._conv4_index(x.device))
            w = conv4.weight.permute(0, 2, 3, 1).reshape(256, -1).to(cols.dtype)
            if fast and cols.dtype == torch.bfloat16:
                y = _c1._LinearReLU.apply(cols.reshape(B * 9, -1), w, conv4.bias.to(cols.dtype))   # [B*9, 256], ReLU applied
            else:
                y = torch.relu(F.linear(cols.reshape(B * 9, -1), w, conv4.bias.to(cols.dtype)))
            if fast and y.dtype == torch.bfloat16:
                # Flatten of [B, 256, 3, 3] orders the features (c, o); y is (o, c): permute fc0's 0.6 M weights instead
                # of transposing the activations (and their gradient) of every sample
                w0 = self.fc0.weight.view(256, 256, 9).permute(0, 2, 1).reshape(256, 2304)
                x = _c1._LinearReLU.apply(y.view(B, 2304), w0.to(torch.bfloat16), self.fc0.bias.to(torch.bfloat16))
            else:
                x = y.view(B, 9, 256).transpose(1, 2).reshape(B, 2304)             # Flatten of [B, 256, 3, 3]
                x = torch.relu(self.fc0(x))
        x = torch.cat([x, position_goal.to(x.dtype)], 1)
        if fast and x.dtype == torch.bfloat16:
            return _c1.linear_relu(x, self.fc1)
        return torch.relu(self.fc1(x))


class Net_PPO_actor(nn.Module):  # all_net.py:191-218
    def __init__(self):
        super().__init__()
        self.bone1 = TINet()
        self.A = nn.Linear(512, 5)
        self.apply(_weights_init)

    def forward(self, state_matrix, position, goal, model="actor"):
        return torch.softmax(self.A(self.bone1(state_matrix, position, goal)).float(), dim=1)

    def logits(self, state_matrix, position, goal):
        return self.A(self.bone1(state_matrix, position, goal)).float()


class Net_PPO_critic(nn.Module):  # all_net.py:222-247
    def __init__(self):
        super().__init__()
        self.bone2 = TINet()
        self.V = nn.Linear(512, 1)
        self.apply(_weights_init)

    def forward(self, state_matrix, position, goal, model="actor"):
        return self.V(self.bone2(state_matrix, position, goal)).float()


_lut_cache = {}


def decode_matrix(codes: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """uint8 featuriser codes -> the float matrix_env values (the LUT is applied here, in the
    network's loader, so the rollout buffer keeps 1 byte per cell)."""
    key = (str(codes.device), dtype)
    if key not in _lut_cache:  # built once per device (a host-to-device copy is not capturable in a CUDA graph)
        _lut_cache[key] = torch.tensor(MATRIX_LUT, dtype=dtype, device=codes.device)
    return _lut_cache[key][codes.long()]


class RolloutBuffer:
    """Device-resident, time-major form of Buffer_gridworld's record (train_ppo.py:93-97):
    s uint8 codes [T,N,5,289] (decoded on load), a int64 [T,N], p float32 [T,N,5,2],
    g float32 [N,2] (constant goal), r, d, a_logp float32 [T,N]."""

    def __init__(self, T: int, N: int, device):
        self.T, self.N, self.device = T, N, device
        self.s = torch.empty((T, N, 5, 289), dtype=torch.uint8, device=device)
        self.p = torch.empty((T, N, 5, 2), dtype=torch.float32, device=device)
        self.a = torch.empty((T, N), dtype=torch.int64, device=device)
        self.r = torch.empty((T, N), dtype=torch.float32, device=device)
        self.d = torch.empty((T, N), dtype=torch.float32, device=device)
        self.a_logp = torch.empty((T, N), dtype=torch.float32, device=device)
        self.g = torch.empty((N, 2), dtype=torch.float32, device=device)
        self.ended = torch.zeros((T, N), dtype=torch.uint8, device=device)   # terminated | truncated
        self.counter = 0

    @property
    def full(self):
        return self.counter >= self.T

    def store(self, s_codes, a, p, r, d, a_logp):
        t = self.counter
        self.s[t].copy_(s_codes); self.a[t].copy_(a); self.p[t].copy_(p)
        self.r[t].copy_(r); self.d[t].copy_(d); self.a_logp[t].copy_(a_logp)
        self.counter += 1

    def flat(self):
        """Sample-major views ([T*N, ...]) in the field names the reference uses."""
        B = self.counter * self.N
        T = self.counter
        # next_same: sample i + N is the NEXT record of the same env and episode, i.e. its frames 0..3 / positions 0..3 are this
        # record's frames 1..4 (VecRollout pushes record t + 1 from record t unless the episode ended at step t), so
        # V(s') of sample i is V(s) of sample i + N and PPO.values evaluates the critic on it once
        nxt = torch.zeros((T, self.N), dtype=torch.bool, device=self.device)
        if T > 1:
            nxt[:T - 1] = self.ended[:T - 1] == 0
        return {"s": self.s[:T].reshape(B, 5, 289), "p": self.p[:T].reshape(B, 5, 2), "a": self.a[:T].reshape(B, 1),
                "g": self.g.unsqueeze(0).expand(T, self.N, 2).reshape(B, 2), "r": self.r[:T].reshape(B, 1),
                "d": self.d[:T].reshape(B, 1), "a_logp": self.a_logp[:T].reshape(B, 1),
                "next_same": nxt.reshape(B), "next_stride": self.N}


def with_her(buf: "RolloutBuffer", seed: int = 9981, env_id0: int = 0, first: int = 0):
    """The rollout's samples followed by their hindsight relabels (train_ppo.py:128-134 with
    args.her on): a dict for PPO.update in which `src` maps every sample to the record it copies.
    first = 4 gives pre_her_func's selection for the PPO + predictor loop (train_ppo_predictor.py:174-180)."""
    from . import her
    flat = buf.flat()
    T, N = buf.counter, buf.N
    extra = her.relabel(buf.p[:T], buf.r[:T], buf.ended[:T], seed, env_id0, first=first)
    base = torch.arange(T * N, device=buf.device)
    out = dict(flat)
    out["src"] = torch.cat([base, extra["src"]])
    out["g"] = torch.cat([flat["g"], extra["g"]])
    out["r"] = torch.cat([flat["r"], extra["r"].view(-1, 1)])
    out["d"] = torch.cat([flat["d"], extra["d"].view(-1, 1)])
    return out


class PPO:
    """soa/agent/PPO.py:41-161 with the same defaults.  Differences, all opt-in or structural:
    no TensorBoard writer / heat-map side effects (out of scope), `select_action` is batched,
    `update` accepts a dict of device tensors, an optional `minibatch` override, bf16 autocast
    and a distributed gradient all-reduce."""

    def __init__(self, device="cpu", autocast: Optional[bool] = None, flat_grads: bool = True, _nets=None):
        self.device = torch.device(device)
        if _nets is None:
            _nets = (Net_PPO_actor(), Net_PPO_critic())   # PPO.py:46-47
        self.actor = _nets[0].to(self.device)
        self.critic = _nets[1].to(self.device)
        self.gamma = 0.99
        self.lr = 0.0001
        self.weight_decay = 0.0001
        self.lr_step_size = 200
        self.lr_gamma = 0.8
        self.batch_size = 128
        self.clip_param = 0.1
        self.K_epochs = 10
        self.entropy_coef = 0.01
        self.use_grad_clip = False
        self.use_lr_decay = False
        self.update_count = 0
        self.max_steps = 0
        self.name = None
        self.autocast = (self.device.type == "cuda") if autocast is None else bool(autocast)
        if self.device.type == "cuda":
            self.actor.to(memory_format=torch.channels_last)
            self.critic.to(memory_format=torch.channels_last)
        # one contiguous gradient buffer per network: a single all-reduce per optimiser step
        self._flat = {}
        if flat_grads:
            for name, net in (("actor", self.actor), ("critic", self.critic)):
                self._flat[name] = self._flatten_grads(net)
        # one multi-tensor kernel per step on the GPU; capturable: the step counter lives on the device
        fused = {"fused": True, "capturable": True} if self.device.type == "cuda" else {}
        self.optimizer_actor = torch.optim.Adam(self.actor.parameters(), lr=self.lr, eps=1e-5, **fused)    # PPO.py:57-58
        self.optimizer_critic = torch.optim.Adam(self.critic.parameters(), lr=self.lr, eps=1e-5, **fused)
        self.scheduler_actor = torch.optim.lr_scheduler.StepLR(self.optimizer_actor, self.lr_step_size, self.lr_gamma)
        self.scheduler_critic = torch.optim.lr_scheduler.StepLR(self.optimizer_critic, self.lr_step_size, self.lr_gamma)
        self.two_streams = self.device.type == "cuda"   # actor / critic passes of update() on two CUDA streams
        self.use_graph = self.device.type == "cuda"     # replay the optimiser step from a CUDA graph
        # capture the gradient all-reduce with the step when there are several ranks (TA_PPO_GRAPH_NCCL=0: eager steps).
        # 2 GPUs, 4096-sample minibatches: 3.8 ms per step from the graph, 6.1 ms eager (the step is ~450 launches)
        self.graph_with_nccl = os.environ.get("TA_PPO_GRAPH_NCCL", "1") == "1"
        self._streams = None
        # the hand-scheduled optimiser step (fused_step.py: explicit kernel list, no autograd, flat fp32 master + bf16
        # shadow + own Adam): built on the first update() of a plain PPO agent on a GPU under bf16 autocast
        self.fused_step = os.environ.get("TA_PPO_FUSED_STEP", "1") == "1"
        self._fused = None
        # the captured optimiser-step graph is kept across update() calls while it stays valid (same buffer storage, sizes
        # and hyper-parameters): the small per-update tensors it reads are persistent copies (TA_PPO_KEEP_GRAPH=0: re-capture
        # in every update, as before)
        self.keep_graph = os.environ.get("TA_PPO_KEEP_GRAPH", "1") == "1"
        # V(s') of a sample whose next record belongs to the same episode is V(s) of that record (PPO.values): one critic
        # pass over the buffer instead of two (TA_PPO_SHARE_NEXT_VALUE=0: two full passes as in the reference)
        self.share_next_value = os.environ.get("TA_PPO_SHARE_NEXT_VALUE", "1") == "1"
        self._static = None
        self._graph_cache = None
        self._step_key = None
        self.last_action_loss = float("nan")
        self.last_value_loss = float("nan")

    @staticmethod
    def _flatten_grads(net):
        params = [p for p in net.parameters()]
        flat = torch.zeros(sum(p.numel() for p in params), dtype=params[0].dtype, device=params[0].device)
        off = 0
        for p in params:
            n = p.numel()
            g = flat[off:off + n].view(p.shape)
            if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
                g = flat[off:off + n].view(p.shape[0], p.shape[2], p.shape[3], p.shape[1]).permute(0, 3, 1, 2)
            p.grad = g
            off += n
        return flat

    # ------------------------------------------------------------------ acting
    def _net_in(self, frames, rows=None, lo=None):
        """What the actor / critic receive for 4 frames [B,4,289] (subclasses add predicted frames).  During update()
        `rows` are the buffer rows the frames were gathered from and `lo` the first of the record's frames used (0 or 1),
        so that a subclass can reuse what it computed per buffer row in _begin_update."""
        return frames

    def _begin_update(self, s):
        """Called by update() once the records `s` [B,5,289] are on the device, before the critic passes."""

    def _end_update(self):
        """Called when update() has issued its last optimiser step."""

    def _amp(self):
        return torch.autocast(device_type=self.device.type, dtype=torch.bfloat16, enabled=self.autocast)

    @torch.no_grad()
    def select_action(self, state_matrix, states_stack, goal, device=None):
        """PPO.py:73-92 for N envs: state_matrix [N,5,289] (float, or uint8 codes), states_stack
        [N,5,2], goal [N,2] -> (action index int64 [N], log-prob float32 [N]).  Uses frames
        1..4 (the newest four) like the reference."""
        return self.select_action_frames(state_matrix[:, 1:5], states_stack[:, 1:5], goal)

    @torch.no_grad()
    def select_action_frames(self, frames, positions, goal):
        """The same on the four current frames directly: frames [N,4,289] (float or uint8 codes),
        positions [N,4,2], goal [N,2]."""
        if frames.dtype == torch.uint8 and not (frames.is_cuda and self.autocast):
            frames = decode_matrix(frames)
        self.actor.eval()
        self.critic.eval()
        x = self._net_in(frames if frames.dtype == torch.uint8 else frames.float())
        with self._amp():
            a_prob = self.actor(x, positions.float(), goal.float())
        # (argument validation reads the probabilities back on the host every step: off on the GPU, where it would also
        # make the rollout impossible to capture in a CUDA graph)
        dist = Categorical(probs=a_prob, validate_args=None if a_prob.device.type != "cuda" else False)
        a = dist.sample()
        return a, dist.log_prob(a)

    # ------------------------------------------------------------------ learning
    @torch.no_grad()
    def values(self, s, p, g, chunk: int = 16384, src=None, next_same=None, next_stride: int = 0):
        """(V(s[:,0:4]), V(s[:,1:5])) in chunks: the two critic passes of PPO.py:113-114.
        next_same (bool [B], with next_stride): where set, sample i + next_stride holds the state s' of sample i as ITS s
        (RolloutBuffer.flat), so V(s') is read from the first pass and the second pass only runs on the remaining samples
        (episode ends and the last time step): the same values, about half the critic work."""
        self.critic.eval()
        B = s.shape[0] if src is None else src.shape[0]

        def critic_pass(lo, rows_of, idx=None):
            outs = []
            n = B if idx is None else idx.numel()
            for i in range(0, n, chunk):
                rows = rows_of(i) if idx is None else idx[i:i + chunk]
                sc = s[rows]
                if sc.dtype == torch.uint8 and not (sc.is_cuda and self.autocast):
                    sc = decode_matrix(sc)
                pc, gc = p[rows], (g[i:i + chunk] if idx is None else g[rows])
                with self._amp():
                    outs.append(self.critic(self._net_in(sc[:, lo:lo + 4], rows, lo), pc[:, lo:lo + 4], gc))
            return torch.cat(outs) if outs else torch.empty((0, 1), device=s.device)

        rows_of = (lambda i: slice(i, i + chunk)) if src is None else (lambda i: src[i:i + chunk])
        v = critic_pass(0, rows_of)
        if next_same is None or src is not None or next_stride <= 0:
            return v, critic_pass(1, rows_of)
        nxt = next_same.to(s.device).view(-1)
        rest = torch.nonzero(~nxt).view(-1)                       # (one host synchronisation per update)
        v_next = torch.empty_like(v)
        shifted = torch.roll(v, -int(next_stride), 0)             # v[i + stride]; the wrapped tail is never selected (next_same is 0 there)
        v_next[nxt] = shifted[nxt].to(v_next.dtype)
        if rest.numel():
            v_next[rest] = critic_pass(1, None, idx=rest).to(v_next.dtype)
        return v, v_next

    def advantages(self, r, v, v_next):
        """PPO.py:112-115: target_v = r + gamma*V(s'), adv = target_v - V(s).  On a GPU this is
        the ta_gae kernel in reference mode; on the CPU (tests, the reference arm) the two
        torch ops the reference uses."""
        if r.is_cuda:
            from . import advantage
            adv, target_v = advantage.reference_mode(r, v, v_next, self.gamma)
            return target_v, adv
        target_v = r + self.gamma * v_next
        return target_v, target_v - v

    def _allreduce(self, name, net, group):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        ws = dist.get_world_size(group)
        if name in self._flat:
            probe = getattr(self, "_probe_events", None)
            if probe is not None:   # probe_step(): device time of the collective alone
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            dist.all_reduce(self._flat[name], group=group)
            if probe is not None:
                e1.record()
                probe.append((name, e0, e1))
            self._flat[name].div_(ws)
        else:
            for p in net.parameters():
                dist.all_reduce(p.grad, group=group)
                p.grad.div_(ws)

    def _make_step(self, buffer, minibatch: Optional[int] = None, group=None):
        """Moves the buffer to the device, runs the two critic passes and the advantage kernel
        (PPO.py:105-115) and returns (step, B, bs, src): step(idx) is one optimiser step on the
        minibatch idx (PPO.py:124-144) and returns the two detached losses."""
        if not isinstance(buffer, dict):
            buffer = {k: torch.as_tensor(buffer[k]) for k in ("s", "p", "a", "g", "r", "a_logp")}
        dev = self.device
        s = buffer["s"].to(dev)
        p = buffer["p"].to(dev, torch.float32)
        a = buffer["a"].to(dev, torch.int64)
        g = buffer["g"].to(dev, torch.float32)
        r = buffer["r"].to(dev, torch.float32).view(-1, 1)
        old_a_logp = buffer["a_logp"].to(dev, torch.float32).view(-1, 1)
        if s.dtype != torch.uint8:
            s = s.to(torch.float32)
        # hindsight-relabelled samples (her.relabel): record i copies s, p, a, a_logp of record src[i]
        # and carries its own g, r -- the fields her_func overrides (env_buffer.py:118-127)
        src = buffer.get("src")
        if src is not None:
            src = src.to(dev)
            assert g.shape[0] == src.shape[0] and r.shape[0] == src.shape[0]
            old_a_logp, a = old_a_logp[src], a[src]
        B = s.shape[0] if src is None else src.shape[0]
        self._begin_update(s)
        v, v_next = self.values(s, p, g, src=src, next_same=buffer.get("next_same") if self.share_next_value else None,
                                next_stride=int(buffer.get("next_stride", 0)))
        target_v, adv = self.advantages(r, v, v_next)
        bs = minibatch or self.batch_size
        self.actor.train()
        self.critic.train()

        streams = None
        if dev.type == "cuda" and self.two_streams:
            if self._streams is None:
                self._streams = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
            streams = self._streams

        fused = self._fused_nets() if (s.dtype == torch.uint8 and s.is_contiguous() and p.is_contiguous()) else None
        if fused is not None:
            step = self._fused_step_fn(fused, s, p, g, a, old_a_logp, adv, target_v, src, streams, group)
            return step, B, bs, src

        def step(idx):
            """One optimiser step on the minibatch `idx` (PPO.py:124-144); returns the two losses."""
            rows = idx if src is None else src[idx]
            sb = s[rows]
            if sb.dtype == torch.uint8 and not (sb.is_cuda and self.autocast):
                sb = decode_matrix(sb)
            sb = self._net_in(sb[:, 0:4], rows, 0)
            pb, gb = p[rows][:, 0:4], g[idx]
            for name in self._flat:
                self._flat[name].zero_()
            if not self._flat:
                self.optimizer_actor.zero_grad()
                self.optimizer_critic.zero_grad()

            def actor_part():
                with self._amp():
                    probs = self.actor(sb, pb, gb)
                # (argument validation reads the probabilities back on the host: not capturable, off on the GPU)
                dist = Categorical(probs=probs, validate_args=None if dev.type != "cuda" else False)
                dist_entropy = dist.entropy().view(-1, 1)
                a_logp = dist.log_prob(a[idx].squeeze(-1)).view(-1, 1)
                ratio = torch.exp(a_logp - old_a_logp[idx])
                surr1 = ratio * adv[idx]
                surr2 = torch.clamp(ratio, 1.0 - self.clip_param, 1.0 + self.clip_param) * adv[idx]
                loss = (-torch.min(surr1, surr2) - self.entropy_coef * dist_entropy).mean()
                loss.backward()
                # this network's gradient all-reduce and optimiser step follow its backward on ITS stream: with two
                # streams the collective overlaps the other network's backward
                self._allreduce("actor", self.actor, group)
                if self.use_grad_clip:
                    torch.nn.utils.clip_grad_norm_(self.actor.parameters(), 0.5)
                self.optimizer_actor.step()
                return loss.detach()

            def critic_part():
                with self._amp():
                    vpred = self.critic(sb, pb, gb)
                loss = F.smooth_l1_loss(vpred, target_v[idx])
                loss.backward()
                self._allreduce("critic", self.critic, group)
                if self.use_grad_clip:
                    torch.nn.utils.clip_grad_norm_(self.critic.parameters(), 0.5)
                self.optimizer_critic.step()
                return loss.detach()

            if streams is None:
                action_loss, value_loss = actor_part(), critic_part()
            else:
                # the two networks are independent: their forward / backward / all-reduce / Adam run on two
                # streams so that one net's small kernels and its collective fill the gaps of the other's (the
                # reference's order of operations inside each network is unchanged)
                cur = torch.cuda.current_stream(dev)
                for st in streams:
                    st.wait_stream(cur)
                with torch.cuda.stream(streams[0]):
                    action_loss = actor_part()
                with torch.cuda.stream(streams[1]):
                    value_loss = critic_part()
                for st in streams:
                    cur.wait_stream(st)
            return action_loss, value_loss

        return step, B, bs, src

    def probe_step(self, buffer, minibatch: Optional[int] = None, group=None):
        """Measurement aid (bench.py, outside every timed region): ONE eager optimiser step on the first minibatch of
        `buffer`, instrumented -- kernels launched by the step (torch profiler) and device time of each network's
        gradient all-reduce (CUDA events around the collective)."""
        if self.device.type != "cuda":
            return {}
        bs = minibatch or self.batch_size
        sub = {k: (v[:bs] if torch.is_tensor(v) else v) for k, v in buffer.items()}
        step, B, bs, _ = self._make_step(sub, bs, group)
        idx = torch.arange(min(B, bs), device=self.device)
        step(idx)
        torch.cuda.synchronize(self.device)
        out = {}
        self._probe_events = None
        try:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                step(idx)
                torch.cuda.synchronize(self.device)
            kernels = [e for e in prof.events() if str(e.device_type).endswith("CUDA") and not e.name.lower().startswith("memcpy")
                       and not e.name.lower().startswith("memset")]
            out["launches_per_optimizer_step"] = len(kernels)
            own = [e for e in kernels if "ta::" in e.name]
            out["own_launches_per_optimizer_step"] = len(own)
            out["kernel_time_per_optimizer_step_us"] = float(sum(e.device_time for e in kernels))
        except Exception as exc:  # noqa: BLE001 -- CUPTI not available: the count is skipped, not guessed
            out["launches_per_optimizer_step"] = None
            out["profiler_error"] = f"{type(exc).__name__}: {exc}"[:200]
            step(idx)
            torch.cuda.synchronize(self.device)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            # each gradient bucket's collective timed alone with CUDA events, in a step of its own that all ranks enter
            # together (inside the profiled step the ranks' CUPTI start-up skews them by tens of milliseconds, and a
            # collective's time includes the wait for the slowest rank)
            dist.barrier(group=group)
            torch.cuda.synchronize(self.device)
            self._probe_events = []
            step(idx)
            torch.cuda.synchronize(self.device)
            if self._probe_events:
                out["allreduce_us_per_optimizer_step"] = {n: e0.elapsed_time(e1) * 1e3 for n, e0, e1 in self._probe_events}
            self._probe_events = None
        if self._fused is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            # the collective on its own (nothing else on the GPU): one network's whole flat gradient buffer, best of 5
            res = {}
            for name, fn in self._fused.items():
                best = None
                for _ in range(5):
                    dist.barrier(group=group)
                    torch.cuda.synchronize(self.device)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    dist.all_reduce(fn.G32, group=group)
                    e1.record()
                    torch.cuda.synchronize(self.device)
                    us = e0.elapsed_time(e1) * 1e3
                    best = us if best is None else min(best, us)
                res[name] = best
            out["allreduce_us_alone"] = dict(res, bytes_per_network=int(self._fused["actor"].G32.numel() * 4),
                                             note="in the step it runs in two buckets overlapped with the convolution stem's backward")
        return out

    def _fused_nets(self):
        """The two networks under the hand-scheduled step (fused_step.FusedNet), or None where it does not apply."""
        from . import fused_step
        if not (self.fused_step and fused_step.supported(self)):
            return None
        if self._fused is None:
            self._fused = {"actor": fused_step.make(self, self.actor, "actor", self.lr, 1e-5),
                           "critic": fused_step.make(self, self.critic, "critic", self.lr, 1e-5)}
            for name, fn in self._fused.items():      # the all-reduce operand is the fused net's flat gradient buffer
                self._flat[name] = fn.G32
                opt = self.optimizer_actor if name == "actor" else self.optimizer_critic
                fn.import_adam_state(opt.state_dict())
        return self._fused

    def _fused_step_fn(self, fused, s, p, g, a, old_a_logp, adv, target_v, src, streams, group):
        """step(idx) through fused_step.FusedNet: gather kernel, then per network (on its own stream) forward, loss,
        backward, gradient all-reduce, Adam + bf16 shadow refresh."""
        import ctypes as C
        from . import _capi
        L, dev = _capi.lib(), self.device
        fa, fc = fused["actor"], fused["critic"]
        fa.lr = float(self.optimizer_actor.param_groups[0]["lr"])
        fc.lr = float(self.optimizer_critic.param_groups[0]["lr"])
        a = a.contiguous(); old_a_logp = old_a_logp.contiguous(); adv = adv.contiguous(); target_v = target_v.contiguous(); g = g.contiguous()
        clip, ent = float(self.clip_param), float(self.entropy_coef)
        ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size(group)
        self._step_key = None
        pred_table = self._pred_table if (fa.IN_CH == 8 and getattr(self, "_pred_valid", False)) else None
        assert (fa.IN_CH == 8) == (pred_table is not None)
        if self.keep_graph and self.use_graph and src is None:
            # everything the step reads besides the (large, caller-owned) frame codes `s` moves into persistent tensors, so
            # that a graph captured over them can be replayed by later update() calls on the same rollout buffer
            small = {"p": p, "g": g, "a": a, "old_a_logp": old_a_logp, "adv": adv, "target_v": target_v}
            sig = tuple((k, tuple(t.shape), t.dtype) for k, t in small.items())
            if self._static is None or self._static["sig"] != sig:
                self._static = {"sig": sig, "t": {k: torch.empty_like(t, memory_format=torch.contiguous_format) for k, t in small.items()}}
                self._graph_cache = None
            for k, t in small.items():
                self._static["t"][k].copy_(t)
            st_t = self._static["t"]
            p, g, a, old_a_logp, adv, target_v = (st_t[k] for k in ("p", "g", "a", "old_a_logp", "adv", "target_v"))
            self._step_key = (sig, s.data_ptr(), tuple(s.shape), tuple(s.stride()), world, id(group), fa.lr, fc.lr, clip, ent,
                              id(fa), id(fc), fa.tc_dgrad, fa.stem_bwd_fused, None if streams is None else tuple(id(x) for x in streams),
                              None if pred_table is None else pred_table.data_ptr())

        def step(idx):
            bs = idx.numel()
            st = lambda: C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            sb = torch.empty((bs, 4, 289), dtype=torch.uint8, device=dev)
            pg16 = torch.empty((bs, 16), dtype=torch.bfloat16, device=dev)
            a_mb = torch.empty(bs, dtype=torch.int32, device=dev)
            old_mb = torch.empty(bs, dtype=torch.float32, device=dev)
            adv_mb = torch.empty(bs, dtype=torch.float32, device=dev)
            tv_mb = torch.empty(bs, dtype=torch.float32, device=dev)
            _capi.check(L.ta_gather_minibatch(ptr(s), ptr(p), ptr(g), ptr(a), ptr(old_a_logp), ptr(adv), ptr(target_v), ptr(idx), ptr(src), bs,
                                              ptr(sb), ptr(pg16), ptr(a_mb), ptr(old_mb), ptr(adv_mb), ptr(tv_mb), st()), "ta_gather_minibatch")

            extra = None
            if pred_table is not None:   # the predictor agent: the 4 predicted frames of the gathered records (fused_step.FusedNet8)
                extra = pred_table.index_select(0, idx if src is None else src[idx]).float()

            def actor_loss(out, d_out, db_head):
                _capi.check(L.ta_ppo_actor_loss(ptr(out), ptr(a_mb), ptr(old_mb), ptr(adv_mb), bs, clip, ent, ptr(d_out), ptr(fa.loss),
                                                ptr(db_head), ptr(fa.step_t), st()), "ta_ppo_actor_loss")

            def critic_loss(out, d_out, db_head):
                _capi.check(L.ta_ppo_critic_loss(ptr(out), ptr(tv_mb), bs, ptr(d_out), ptr(fc.loss), ptr(db_head), ptr(fc.step_t), st()),
                            "ta_ppo_critic_loss")

            def part(fn, name, net, loss_fn):
                # gradient all-reduce in two buckets on this network's stream: the late layers' (91 % of the bytes) is
                # issued asynchronously before the stem's backward and waited for with the small stem bucket at the end
                works = []

                def reduce(t, last):
                    probe = getattr(self, "_probe_events", None)
                    if probe is not None:   # probe_step(): device time of each bucket's collective, not overlapped with anything
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        torch.distributed.all_reduce(t, group=group)
                        e1.record()
                        probe.append((name + ("/stem" if last else "/late_layers"), e0, e1))
                        return
                    works.append(torch.distributed.all_reduce(t, group=group, async_op=True))
                    if last:
                        for w in works:
                            w.wait()

                fn.forward_backward(sb, pg16, loss_fn, reduce if world > 1 else None, extra)
                fn.adam(1.0 / world)          # the SUM's 1 / world is folded into Adam's gradient read
                return fn.loss

            if streams is None:
                la, lc = part(fa, "actor", self.actor, actor_loss), part(fc, "critic", self.critic, critic_loss)
            else:
                cur = torch.cuda.current_stream(dev)
                for sx in streams:
                    sx.wait_stream(cur)
                with torch.cuda.stream(streams[0]):
                    la = part(fa, "actor", self.actor, actor_loss)
                with torch.cuda.stream(streams[1]):
                    lc = part(fc, "critic", self.critic, critic_loss)
                for sx in streams:
                    cur.wait_stream(sx)
                for t in (sb, pg16, a_mb, old_mb, adv_mb, tv_mb, idx) + (() if extra is None else (extra,)):   # read on the side streams
                    t.record_stream(streams[0]); t.record_stream(streams[1])
            return la, lc

        return step

    def _common_steps(self, B: int, bs: int, group) -> int:
        """Optimiser steps per epoch, agreed between ranks: every rank issues one all-reduce per step, so all of
        them must run the same number of steps.  With equal B (the plain rollout) this is ceil(B / bs) as in the
        reference; when B differs between ranks (hindsight relabels append a data-dependent number of samples per
        rank) every rank runs min over ranks of B // bs full minibatches of its own permutation and drops the rest."""
        import torch.distributed as dist
        n_steps = (B + bs - 1) // bs
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return n_steps
        t = torch.tensor([B, -B], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        b_min, b_max = int(t[0].item()), -int(t[1].item())
        if b_min == b_max:
            return n_steps
        return max(1, b_min // bs)

    def update(self, buffer, device=None, i_ep: int = 0, minibatch: Optional[int] = None, epochs: Optional[int] = None,
               group=None, sampler_generator=None):
        """PPO.py:103-158.  `buffer` is a dict of tensors with the reference's field names
        (s [B,5,289] float or uint8 codes, p [B,5,2], a [B,1], g [B,2], r [B,1], a_logp [B,1]) or
        a numpy structured array like Buffer_gridworld.buffer."""
        dev = self.device
        timing = os.environ.get("TA_PPO_TIMING", "0") == "1" and dev.type == "cuda"   # measurement aid: synchronising wall-clock marks

        def mark(name, _t=[None]):
            if timing:
                import time
                torch.cuda.synchronize(dev)
                now = time.perf_counter()
                if _t[0] is not None:
                    self.last_update_phases[name] = self.last_update_phases.get(name, 0.0) + (now - _t[0]) * 1e3
                _t[0] = now

        self.last_update_phases = {}
        mark("start")
        step, B, bs, src = self._make_step(buffer, minibatch, group)
        n_steps = self._common_steps(B, bs, group)
        mark("prepare_ms")

        def minibatches():
            # PPO.py:122: BatchSampler(SubsetRandomSampler(range(len(buffer))), batch_size, drop_last=False).
            # On the CPU that exact sampler (same torch RNG stream as the reference); on the GPU the
            # same thing -- a fresh random permutation per epoch cut into batches -- drawn on the device
            # (a Python-level sampler over millions of indices would cost more than the update itself).
            if dev.type != "cuda" or sampler_generator is not None:
                for k, sample_index in enumerate(BatchSampler(SubsetRandomSampler(range(B), generator=sampler_generator),
                                                              batch_size=bs, drop_last=False)):
                    if k < n_steps:
                        yield torch.as_tensor(sample_index, device=dev)
            else:
                perm = torch.randperm(B, device=dev)
                for k in range(n_steps):
                    yield perm[k * bs:(k + 1) * bs]

        # The step is ~300 small launches; driven from Python the update is CPU-bound.  On one GPU the
        # whole step (gather, both nets forward / backward, Adam) is captured ONCE per update() into a
        # CUDA graph over a static index buffer and replayed for every full minibatch (same steps,
        # same order as the eager loop).
        world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size(group)
        # (with several ranks the gradient all-reduce is captured too: every rank replays the same graph)
        want_graph = dev.type == "cuda" and self.use_graph and (world == 1 or self.graph_with_nccl) and bool(self._flat)
        graph, idx_static, loss_static, eager_full = None, None, None, 0
        key = (self._step_key, bs) if (want_graph and self._step_key is not None) else None
        if key is not None and self._graph_cache is not None and self._graph_cache["key"] == key:
            graph, idx_static, loss_static = (self._graph_cache[k] for k in ("graph", "idx", "loss"))
        elif self._graph_cache is not None:
            self._graph_cache = None     # (frees the stale graph and its memory pool)
        ev_first = ev_last = None
        n_replayed = 0
        side = torch.cuda.Stream(device=dev) if want_graph else None

        self._last = None
        for _ in range(epochs or self.K_epochs):
            for idx in minibatches():
                full = idx.numel() == bs
                if want_graph and full and graph is None and eager_full >= 2:
                    # two eager steps have warmed the allocator and the cuDNN / cuBLAS plans up: capture the
                    # step once (capture records, it does not execute) and replay it from here on
                    idx_static = idx.clone()
                    torch.cuda.synchronize(dev)
                    mark("eager_steps_ms")
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        loss_static = step(idx_static)
                    mark("capture_ms")
                if graph is not None and full:
                    if ev_first is None:
                        ev_first, ev_last = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        ev_first.record()
                    idx_static.copy_(idx)
                    graph.replay()
                    n_replayed += 1
                    self._last = loss_static
                elif want_graph and full:
                    mark("sampler_ms")
                    side.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(side):   # warm-up steps run on a side stream, as capture will
                        self._last = step(idx)
                    torch.cuda.current_stream(dev).wait_stream(side)
                    eager_full += 1
                    mark("eager_step%d_ms" % eager_full)
                else:
                    self._last = step(idx)
                self.update_count += 1
        if graph is not None:
            if ev_first is not None:
                ev_last.record()
            torch.cuda.synchronize(dev)
            if ev_first is not None:   # device time of the graph-replayed optimiser steps alone (bench.py reports it)
                self.last_replay_stats = {"steps": n_replayed, "ms_per_step": ev_first.elapsed_time(ev_last) / max(1, n_replayed)}
            self._last = tuple(x.clone() for x in self._last)
            mark("replay_ms")
            if key is not None:
                self._graph_cache = {"key": key, "graph": graph, "idx": idx_static, "loss": loss_static}
            del graph
            mark("graph_free_ms")
        self.last_action_loss, self.last_value_loss = (float(x) for x in self._last)
        self._end_update()
        if dev.type == "cuda":
            # the tcgen05 kernels bound their MMA-barrier waits and raise a flag instead of hanging: fail loudly here
            from . import _capi
            if _capi.lib().ta_debug_conv1_tc_failed() != 0:
                raise RuntimeError("a tcgen05 conv1 kernel gave up waiting for its MMA barrier: the results of this update are invalid")
        if self.use_lr_decay:
            self.scheduler_actor.step()
            self.scheduler_critic.step()
        return self.last_action_loss, self.last_value_loss

    # ------------------------------------------------------------------ checkpoints (PPO.py:94-101)
    def state_dict(self, i_ep: int = 0):
        oa, oc = self.optimizer_actor.state_dict(), self.optimizer_critic.state_dict()
        if self._fused is not None:   # the Adam moments live in the fused nets' flat buffers: same torch.optim.Adam format
            oa = self._fused["actor"].export_adam_state(self.optimizer_actor)
            oc = self._fused["critic"].export_adam_state(self.optimizer_critic)
        return {"model_actor": self.actor.state_dict(), "model_critic": self.critic.state_dict(),
                "optimizer_actor": oa, "optimizer_critic": oc, "epoch": i_ep}

    def load_state_dict(self, state):
        self.actor.load_state_dict(state["model_actor"])
        self.critic.load_state_dict(state["model_critic"])
        if "optimizer_actor" in state:
            self.optimizer_actor.load_state_dict(state["optimizer_actor"])
            self.optimizer_critic.load_state_dict(state["optimizer_critic"])
        if self._fused is not None:
            for name, fn in self._fused.items():
                fn.refresh()
                if "optimizer_" + name in state:
                    fn.import_adam_state(state["optimizer_" + name])

    def save_param(self, path, i_ep: int = 0):
        torch.save(self.state_dict(i_ep), path)

    def broadcast_parameters(self, group=None, src: int = 0):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            for net in (self.actor, self.critic):
                for t in list(net.parameters()) + list(net.buffers()):
                    dist.broadcast(t.data, src, group=group)
            if self._fused is not None:
                for fn in self._fused.values():
                    fn.refresh()


class VecRollout:
    """soa/train_ppo.py:99-160 for N envs at once: Env_transact.reset / env_action / step,
    frame-stack roll and Buffer_gridworld.store, all on the device.  One collect() fills a
    [T, N] RolloutBuffer; episodes end / restart independently per env, and a new episode's
    stack is the tiled reset frame exactly like Env_transact.reset (env_buffer.py:420-423).

    The frame stacks are never rolled in place: record t is written straight into the buffer by
    ta_stack_push from record t-1 (or from the tiled reset frame for envs whose episode ended at
    t-1), so each step moves 4 frames in and 5 frames out per env and nothing else."""

    def __init__(self, env, agent: PPO, horizon: int):
        assert horizon >= 2
        self.env, self.agent, self.T = env, agent, int(horizon)
        N, dev = env.num_envs, env.device
        assert not env.autoreset, "VecRollout drives the resets itself (the terminal frame is part of the record)"
        # record t is written in place at byte offset t * N * 1445 of the buffer; ta_stack_push wants 16-byte aligned
        # records, so the per-step stride must be a multiple of 16 bytes (1445 is odd)
        if env.num_envs % 16:
            raise ValueError(f"VecRollout needs num_envs % 16 == 0 (got {env.num_envs}): the device rollout buffer is written "
                             f"in place by 16-byte vector stores; pad the env count per rank")
        self.buffer = RolloutBuffer(self.T, N, dev)
        self.buffer.g[:] = torch.tensor([float(env.goal_pos[1]), float(env.goal_pos[0])], device=dev)  # data_env: (y, x)
        self.amap = torch.tensor(POLICY_TO_ENV_ACTION, dtype=torch.uint8, device=dev)
        self.ep_return = torch.zeros(N, dtype=torch.float32, device=dev)
        env.reset()
        # the stack Env_transact.reset builds (a constant: _gen_grid is deterministic)
        s0 = torch.empty((N, 5, 289), dtype=torch.uint8, device=dev)
        p0 = torch.empty((N, 5, 2), dtype=torch.float32, device=dev)
        env.stack_push(None, s0, None, p0, init_all=True)
        self.reset_s, self.reset_p = s0[0, 1:5].clone(), p0[0, 1:5].clone()
        self.prev_s = self.prev_p = None
        self.prev_done = torch.ones(N, dtype=torch.uint8, device=dev)
        self._done_buf = torch.zeros(N, dtype=torch.uint8, device=dev)   # persistent: the captured rollout graph reads / writes it
        self._out = {}
        self.use_graph = os.environ.get("TA_ROLLOUT_GRAPH", "1") == "1"
        self._graph = None

    def current_frames(self):
        """The policy's input: the four newest frames / positions of every env."""
        N = self.env.num_envs
        if self.prev_s is None:
            return self.reset_s.expand(N, 4, 289), self.reset_p.expand(N, 4, 2)
        m = self.prev_done.bool()[:, None, None]
        return torch.where(m, self.reset_s, self.prev_s[:, 1:5]), torch.where(m, self.reset_p, self.prev_p[:, 1:5])

    @torch.no_grad()
    def collect(self):
        """T steps of the reference's inner loop (train_ppo.py:108-123) for every env: select
        action, env.step (no autoreset), featurise the post-step state into record t, then
        MiniGridEnv.reset for the envs whose episode ended (what the reference does at the top of
        its next episode, train_ppo.py:104-105).

        On a GPU the whole T-step loop is launch-bound from Python (~110 launches per step), so from the second call on
        it is captured ONCE into a CUDA graph (every buffer it touches is persistent: the rollout buffer, the env
        state, the previous record) and replayed: same kernels, same order, fresh random actions per replay (the
        sampler's Philox offset advances with every replay)."""
        buf = self.buffer
        if self._graph is not None:
            self._graph.replay()
            buf.counter = self.T
            return buf
        if self.use_graph and self.prev_s is not None and type(self.agent) is PPO and self.env.device.type == "cuda":
            dev = self.env.device
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._collect_eager()
            self._graph = g
            g.replay()
            buf.counter = self.T
            return buf
        return self._collect_eager()

    def _collect_eager(self):
        env, buf = self.env, self.buffer
        buf.counter = 0
        for t in range(self.T):
            frames, positions = self.current_frames()
            a_idx, a_logp = self.agent.select_action_frames(frames, positions, buf.g)
            obs, rew, term, trunc, _ = env.step(self.amap[a_idx], out=self._out)
            self._out = {"obs": obs, "reward": rew, "terminated": term.view(torch.uint8), "truncated": trunc.view(torch.uint8)}
            done = (term | trunc).view(torch.uint8)
            env.stack_push(self.prev_s, buf.s[t], self.prev_p, buf.p[t], self.prev_done, init_all=self.prev_s is None)
            buf.a[t].copy_(a_idx); buf.r[t].copy_(rew); buf.d[t].copy_(term); buf.a_logp[t].copy_(a_logp)
            buf.ended[t].copy_(done)
            buf.counter += 1
            self.ep_return += rew
            self.ep_return.masked_fill_(done.bool(), 0.0)
            self._done_buf.copy_(done)
            self.prev_s, self.prev_p, self.prev_done = buf.s[t], buf.p[t], self._done_buf
            env.reset_masked(done)
        return buf
