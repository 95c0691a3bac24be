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
    def __init__(self):
        super().__init__()
        self.extrap_t = 4
        self.nt = 8
        self.recurrent_model = nn.LSTM(1024, 1024, num_layers=3, batch_first=True)
        self.h_0 = torch.zeros(3, 1024)
        self.c_0 = torch.zeros(3, 1024)
        self.device = None

    def forward(self, z_content):
        B, T, D, W, H = z_content.shape
        z_content = z_content.reshape(B, T, D * W * H)
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

    def load_predictor(self, state):
        """train_ppo_predictor.py:77-81: the pre-trained predictor stack from a checkpoint dict."""
        self.encoder.load_state_dict(state["model_encoder"])
        self.decoder.load_state_dict(state["model_decoder"])
        self.predictor.load_state_dict(state["model_predictor"])

    @torch.no_grad()
    def pred_states(self, state_matrix):
        """PPO_Predictor.py:70-83: [B,4,289] current frames -> [B,4,289] predicted next frames."""
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

    def _net_in(self, frames):
        """PPO_Predictor.py:100-103 / :133-139 / :149-150: current frames + the predicted ones."""
        return self._cat(frames)

    def state_dict(self, i_ep: int = 0):
        d = super().state_dict(i_ep)
        d.update(model_encoder=self.encoder.state_dict(), model_decoder=self.decoder.state_dict(),
                 model_predictor=self.predictor.state_dict())
        return d


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
