"""Advantage / returns on the GPU (the ta_gae / ta_adv_* entry points).

reference_mode() is soa/agent/PPO.py:112-115 verbatim (1-step TD target, no done mask, no
normalisation); gae() is the general GAE(gamma, lambda) the same kernel computes.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _capi


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def gae(reward: torch.Tensor, value: torch.Tensor, done: Optional[torch.Tensor] = None, gamma: float = 0.99,
        lam: float = 0.95, use_mask: bool = True, v_next: Optional[torch.Tensor] = None,
        last_value: Optional[torch.Tensor] = None, normalize: bool = False, group=None, out=None):
    """reward/value/done/v_next: [T, N] time-major fp32 (done uint8/bool); last_value: [N].
    Returns (adv, ret) fp32 [T, N].  normalize=True applies (adv-mean)/(std+1e-8) over all
    T*N elements; with a torch.distributed group the three moments are all-reduced first.
    out = (adv, ret): optional preallocated contiguous fp32 [T, N] outputs."""
    assert reward.is_cuda and reward.dtype == torch.float32 and reward.dim() == 2
    T, N = reward.shape
    reward, value = reward.contiguous(), value.contiguous()
    assert value.shape == (T, N) and value.dtype == torch.float32
    if v_next is not None:
        v_next = v_next.contiguous()
        assert v_next.shape == (T, N) and v_next.dtype == torch.float32
    if last_value is not None:
        last_value = last_value.contiguous().view(-1)
        assert last_value.numel() == N and last_value.dtype == torch.float32
    if done is not None:
        done = done.contiguous().view(torch.uint8) if done.dtype == torch.bool else done.to(torch.uint8).contiguous()
    if out is not None:
        adv, ret = out
        for o in (adv, ret):
            assert o.shape == (T, N) and o.dtype == torch.float32 and o.is_contiguous() and o.device == reward.device
    else:
        adv = torch.empty_like(reward)
        ret = torch.empty_like(reward)
    L = _capi.lib()
    st = _stream(reward.device)
    if not normalize:
        _capi.check(L.ta_gae(_ptr(reward), _ptr(value), _ptr(v_next), _ptr(last_value), _ptr(done), float(gamma), float(lam),
                             int(bool(use_mask)), T, N, _ptr(adv), _ptr(ret), st), "ta_gae")
    elif group is None and not (torch.distributed.is_available() and torch.distributed.is_initialized()
                                and torch.distributed.get_world_size() > 1):
        # one rank: the library normalises inside the GAE launch when the whole grid is resident at once (small rollouts),
        # else moments from the GAE launch + one pass over adv
        work = torch.empty(8 + 2 * 1024, dtype=torch.float64, device=reward.device)   # TA_GAE_WORK_DOUBLES
        _capi.check(L.ta_gae_normalized(_ptr(reward), _ptr(value), _ptr(v_next), _ptr(last_value), _ptr(done), float(gamma),
                                        float(lam), int(bool(use_mask)), T, N, _ptr(adv), _ptr(ret), _ptr(work), st),
                    "ta_gae_normalized")
    else:
        # the moments of adv come out of the same launch; normalising is one more pass over adv
        stats = torch.empty(3, dtype=torch.float64, device=reward.device)
        _capi.check(L.ta_gae_stats(_ptr(reward), _ptr(value), _ptr(v_next), _ptr(last_value), _ptr(done), float(gamma), float(lam),
                                   int(bool(use_mask)), T, N, _ptr(adv), _ptr(ret), _ptr(stats), st), "ta_gae_stats")
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                                 and torch.distributed.get_world_size() > 1):
            torch.distributed.all_reduce(stats, group=group)
        _capi.check(L.ta_adv_normalize(_ptr(adv), T * N, _ptr(stats), st), "ta_adv_normalize")
    return adv, ret


def reference_mode(reward, value, v_next, gamma: float = 0.99):
    """target_v = r + gamma*V(s'); adv = target_v - V(s)  (PPO.py:112-115). Inputs of any
    shape with B elements (the reference uses [B,1]); returns (adv, target_v) of that shape."""
    shape = reward.shape
    adv, tv = gae(reward.reshape(1, -1), value.reshape(1, -1), None, gamma, 0.0, use_mask=False,
                  v_next=v_next.reshape(1, -1))
    return adv.view(shape), tv.view(shape)
