"""Hindsight relabelling on the device (ta_her_plan).

Mirrors Buffer_gridworld.her_func (soa/env_buffer.py:101-143), which the reference runs after every
episode while running_score < 0 (soa/train_ppo.py:128-134): up to four distinct positions the agent
visited become goals, and the episode prefix that reached each of them is appended to the buffer
with r = 0.9 / d = 1 on its last record.  Here the copies stay virtual: `relabel` returns, for every
appended record, the index of the rollout record it copies plus the overridden g, r, d -- exactly
the fields her_func changes -- and PPO.update gathers s, p, a, a_logp through `src`.

Buffer_gridworld.pre_her_func (soa/env_buffer.py:145-210), the same thing for the 9-frame records of
soa/train_ppo_predictor.py, is `first=4`: a 9-frame record exists from the episode's 5th step on and
record j is step j + 4, whose frames 0..4 / a[0] / r[0] / a_logp[0] -- all PPO_Predictor.update reads
(PPO_Predictor.py:124-163) -- are the 5-frame record of step j.  The prefix 0..index plus the four
shifted pad records pre_her_func appends (:173-196) are therefore the 5-frame records of steps
0..index+4 with the goal reached at step index+4, and r = 0.9 arrives in r[0] of the 4th pad, i.e. on
step index+4 (tests/golden/make_golden_her.py runs the reference and pins this).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _capi


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def plan(p: torch.Tensor, done: torch.Tensor, seed: int = 9981, env_id0: int = 0, chosen: Optional[torch.Tensor] = None,
         want_unique: bool = False, first: int = 0):
    """p float32 [T,N,5,2], done uint8/bool [T,N] -> plan uint16 [T,N,4] (see include/twoarmy_b200.h).
    With want_unique also returns (indices uint8 [T,N,64], count uint8 [T,N]).  first = 0: her_func,
    first = 4: pre_her_func (record indices stay relative to the episode start)."""
    assert p.is_cuda and p.dtype == torch.float32 and p.dim() == 4 and p.shape[2:] == (5, 2)
    T, N = p.shape[:2]
    p = p.contiguous()
    done = (done.view(torch.uint8) if done.dtype == torch.bool else done.to(torch.uint8)).contiguous()
    assert done.shape == (T, N)
    out = torch.empty((T, N, 4), dtype=torch.uint16, device=p.device)
    uniq = torch.empty((T, N, 64), dtype=torch.uint8, device=p.device) if want_unique else None
    m = torch.empty((T, N), dtype=torch.uint8, device=p.device) if want_unique else None
    if chosen is not None:
        chosen = chosen.to(device=p.device, dtype=torch.uint8).contiguous()
        assert chosen.shape == (T, N, 4)
    st = C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
    _capi.check(_capi.lib().ta_her_plan(_ptr(p), _ptr(done), T, N, int(first), C.c_uint64(seed & (2**64 - 1)), C.c_uint64(env_id0),
                                        _ptr(chosen), _ptr(uniq), _ptr(m), _ptr(out), st), "ta_her_plan")
    return (out, uniq, m) if want_unique else out


def relabel(buf_p: torch.Tensor, buf_r: torch.Tensor, done: torch.Tensor, seed: int = 9981, env_id0: int = 0,
            chosen: Optional[torch.Tensor] = None, first: int = 0):
    """buf_p [T,N,5,2], buf_r [T,N], done [T,N] -> dict(src int64 [M] (flat index t*N+env of the
    copied record), g float32 [M,2] = (y, x), r float32 [M], d float32 [M], slot int64 [M])."""
    pl = plan(buf_p, done, seed, env_id0, chosen, first=first).to(torch.int32)   # [T,N,4]
    T, N = buf_r.shape
    nz = torch.nonzero(pl != 0xFFFF)                                        # [M,3] = (t, env, slot)
    v = pl[nz[:, 0], nz[:, 1], nz[:, 2]]
    src = nz[:, 0] * N + nz[:, 1]
    last = (v & 0x8000) != 0
    g = torch.stack([((v >> 5) & 31).float(), (v & 31).float()], 1)
    r_src = buf_r.reshape(-1)[src]
    r = torch.where(last, torch.full_like(r_src, 0.9), r_src)
    return {"src": src, "g": g, "r": r, "d": last.float(), "slot": nz[:, 2]}
