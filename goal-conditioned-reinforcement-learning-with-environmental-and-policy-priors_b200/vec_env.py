"""TwoarmyVecEnv -- N parallel MiniGrid Twoarmy-17x17 grids on one B200.

Vector form of the reference's gym boundary (gym_minigrid/__init__.py:10-20,
gym_minigrid/envs/twoarmy_v4.py, twoarmy_v6.py, gym_minigrid/minigrid.py:835-1590): same ids,
same reset()/step() meaning, same kwargs (unknown kwargs such as seed= / new_step_api= are
absorbed exactly as MiniGridEnv.__init__(**kwargs) absorbs them, minigrid.py:879).

Everything numeric happens in the CUDA library behind include/twoarmy_b200.h; torch only
owns the output buffers and the stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _capi

ENV_IDS = {"MiniGrid-twoarmy-17x17-v4": 4, "MiniGrid-twoarmy-17x17-v6": 6}
MISSION = "get to the green goal square"  # twoarmy_v4.py:25-27,80
GOAL_POS = (14, 2)                        # twoarmy_v4.py:9
MAX_STEPS = 50                            # twoarmy_v4.py:32
NUM_ACTIONS = 7                           # MiniGridEnv.Actions, minigrid.py:849-864
POLICY_TO_ENV_ACTION = (0, 1, 2, 3, 6)    # Env_transact.env_action, soa/env_buffer.py:364-376

# numpy mirror of `ta_env_state`
STATE_DTYPE = np.dtype(
    [("grid", np.uint8, (289,)), ("agent_x", np.uint8), ("agent_y", np.uint8), ("flags", np.uint8),
     ("risk_count", np.uint8), ("error", np.uint8), ("balls", np.uint8, (10, 2)), ("pad_", np.uint8, (2,)),
     ("step_count", np.int32), ("step_move", np.int32), ("t", np.uint32)], align=True)
assert STATE_DTYPE.itemsize == 328

_ACT_DTYPES = {torch.int32: _capi.ACT_I32, torch.uint8: _capi.ACT_U8, torch.int64: _capi.ACT_I64}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


class TwoarmyVecEnv:
    def __init__(self, env_id="MiniGrid-twoarmy-17x17-v4", num_envs: int = 1, agent_view_size: int = 17,
                 device="cuda:0", seed: int = 9981, env_id0: int = 0, autoreset: bool = True, **kwargs):
        if isinstance(env_id, str):
            if env_id not in ENV_IDS:
                raise KeyError(f"unknown env id {env_id!r}; registered: {sorted(ENV_IDS)}")
            version = ENV_IDS[env_id]
        else:
            version = int(env_id)
        # minigrid.py:903-905
        assert agent_view_size % 2 == 1
        assert agent_view_size >= 3
        if agent_view_size > 17:
            raise NotImplementedError("agent_view_size > 17 (wider than the grid) is not built")
        if not torch.cuda.is_available():
            raise _capi.TwoarmyLibraryError("TwoarmyVecEnv needs a CUDA device: the product has no CPU path")
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.version, self.num_envs, self.view = version, int(num_envs), int(agent_view_size)
        self.autoreset = bool(autoreset)
        self.seed, self.env_id0 = int(seed) & (2**64 - 1), int(env_id0)
        self.mission, self.goal_pos, self.max_steps = MISSION, GOAL_POS, MAX_STEPS
        self.ignored_kwargs = dict(kwargs)  # seed=, new_step_api=, tile_size= ... (train_ppo.py:80-85)
        self._L = _capi.lib()
        h = C.c_void_p()
        _capi.check(self._L.ta_create(C.byref(h), version, self.num_envs, self.view, self.device.index,
                                      C.c_uint64(self.seed), C.c_uint64(self.env_id0)), "ta_create")
        self._h = h
        self._reset_template = None

    # ------------------------------------------------------------------ plumbing
    def close(self):
        if getattr(self, "_h", None):
            self._L.ta_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _new_obs(self):
        return torch.empty((self.num_envs, self.view, self.view, 3), dtype=torch.uint8, device=self.device)

    # ------------------------------------------------------------------ gym-like API
    def reset(self, mask: Optional[torch.Tensor] = None, hard: bool = False) -> torch.Tensor:
        """MiniGridEnv.reset on the masked envs (all when mask is None); returns gen_obs() of
        every env.  Like the reference it leaves the Twoarmy flags alone unless hard=True."""
        if mask is not None:
            mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        obs = self._new_obs()
        _capi.check(self._L.ta_reset(self._h, _ptr(mask), int(hard), _ptr(obs), self._stream()), "ta_reset")
        return obs

    def reset_masked(self, mask: torch.Tensor) -> None:
        """MiniGridEnv.reset on the masked envs without producing observations."""
        mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _capi.check(self._L.ta_reset(self._h, _ptr(mask), 0, None, self._stream()), "ta_reset")

    def observe(self) -> torch.Tensor:
        """gen_obs() of the current state (no transition)."""
        none = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
        obs = self._new_obs()
        _capi.check(self._L.ta_reset(self._h, _ptr(none), 0, _ptr(obs), self._stream()), "ta_reset(observe)")
        return obs

    def observe_general(self, agent_dir=3, see_through_walls=True) -> torch.Tensor:
        """gen_obs() of the current state for any agent_dir (int, or a uint8 tensor [N]) and either visibility
        mode (bool, or a uint8 tensor [N]; False runs Grid.process_vis) -- minigrid.py:1443-1496, 795-832."""
        dirs = stw = None
        if torch.is_tensor(agent_dir):
            dirs = agent_dir.to(device=self.device, dtype=torch.uint8).contiguous()
            assert dirs.numel() == self.num_envs
        if torch.is_tensor(see_through_walls):
            stw = see_through_walls.to(device=self.device, dtype=torch.uint8).contiguous()
            assert stw.numel() == self.num_envs
        obs = self._new_obs()
        _capi.check(self._L.ta_observe_general(self._h, _ptr(dirs), 0 if dirs is not None else int(agent_dir), _ptr(stw),
                                               1 if stw is not None else int(bool(see_through_walls)), _ptr(obs), self._stream()),
                    "ta_observe_general")
        return obs

    def step(self, actions: torch.Tensor, draws: Optional[torch.Tensor] = None, out: Optional[dict] = None,
             want_consumed: bool = False):
        """One env.step for every env.  Returns (obs, reward, terminated, truncated, info).
        obs is the observation env.step itself returns (terminal obs on done steps); with
        autoreset the env is reset afterwards, as the reference's caller does."""
        if actions.device != self.device:
            actions = actions.to(self.device)
        if actions.dtype not in _ACT_DTYPES:
            actions = actions.to(torch.int32)
        actions = actions.contiguous()
        assert actions.numel() == self.num_envs
        n = self.num_envs
        if out is None:
            out = {}
        obs = out.get("obs") if out.get("obs") is not None else self._new_obs()
        rew = out.get("reward") if out.get("reward") is not None else torch.empty(n, dtype=torch.float32, device=self.device)
        term = out.get("terminated") if out.get("terminated") is not None else torch.empty(n, dtype=torch.uint8, device=self.device)
        trunc = out.get("truncated") if out.get("truncated") is not None else torch.empty(n, dtype=torch.uint8, device=self.device)
        cons = None
        if draws is not None:
            draws = draws.to(device=self.device, dtype=torch.uint8).contiguous()
            assert draws.shape == (n, 8)
        if want_consumed or draws is not None:
            cons = torch.empty(n, dtype=torch.uint8, device=self.device)
        flags = _capi.TA_STEP_AUTORESET if self.autoreset else 0
        _capi.check(self._L.ta_step(self._h, _ptr(actions), _ACT_DTYPES[actions.dtype], _ptr(draws), flags, _ptr(obs),
                                    _ptr(rew), _ptr(term), _ptr(trunc), _ptr(cons), self._stream()), "ta_step")
        info = {"consumed": cons} if cons is not None else {}
        return obs, rew, term.view(torch.bool), trunc.view(torch.bool), info

    def step_host(self, actions: np.ndarray, obs: np.ndarray, reward: np.ndarray, terminated: np.ndarray,
                  truncated: np.ndarray, dma: bool = False):
        """End-to-end call with HOST arrays (numpy): H2D actions, fused kernel, D2H results, synchronised on
        return.  Default: the observations cross PCIe as 2-bit codes and host threads expand them into `obs`
        (ta_step_host); dma=True: the expanded bytes are copied straight into `obs` (pin it for speed)."""
        assert actions.dtype in (np.int32, np.uint8, np.int64) and actions.size == self.num_envs
        assert obs.dtype == np.uint8 and obs.size == self.num_envs * 3 * self.view * self.view and obs.flags.c_contiguous
        assert reward.dtype == np.float32 and terminated.dtype == np.uint8 and truncated.dtype == np.uint8
        dt = {np.dtype(np.int32): 0, np.dtype(np.uint8): 1, np.dtype(np.int64): 2}[actions.dtype]
        flags = (_capi.TA_STEP_AUTORESET if self.autoreset else 0) | (_capi.TA_STEP_HOST_DMA if dma else 0)
        _capi.check(self._L.ta_step_host(self._h, C.c_void_p(actions.ctypes.data), dt, flags,
                                         C.c_void_p(obs.ctypes.data), C.c_void_p(reward.ctypes.data),
                                         C.c_void_p(terminated.ctypes.data), C.c_void_p(truncated.ctypes.data)),
                    "ta_step_host")

    def host_d2h_bytes(self) -> int:
        """Device-to-host bytes the last step_host() moved."""
        return int(self._L.ta_step_host_d2h_bytes(self._h))

    def step_packed(self, actions: torch.Tensor):
        """One env.step with the results in transfer form (ta_step_packed): (codes uint32 [npad/32, 2*V*V],
        status uint8 [npad])."""
        if actions.dtype not in _ACT_DTYPES:
            actions = actions.to(torch.int32)
        actions = actions.to(self.device).contiguous()
        ntiles = (self.num_envs + 31) // 32
        codes = torch.empty((ntiles, 2 * self.view * self.view), dtype=torch.int32, device=self.device)
        status = torch.zeros(ntiles * 32, dtype=torch.uint8, device=self.device)
        flags = _capi.TA_STEP_AUTORESET if self.autoreset else 0
        _capi.check(self._L.ta_step_packed(self._h, _ptr(actions), _ACT_DTYPES[actions.dtype], None, flags, _ptr(codes),
                                           _ptr(status), None, self._stream()), "ta_step_packed")
        return codes, status

    def rollout(self, actions: torch.Tensor):
        """T steps with a pre-sampled [T, N] action tensor (autoreset, Philox draws)."""
        assert actions.dim() == 2 and actions.shape[1] == self.num_envs
        if actions.dtype not in _ACT_DTYPES:
            actions = actions.to(torch.int32)
        actions = actions.to(self.device).contiguous()
        T, n = actions.shape
        obs = torch.empty((T, n, self.view, self.view, 3), dtype=torch.uint8, device=self.device)
        rew = torch.empty((T, n), dtype=torch.float32, device=self.device)
        term = torch.empty((T, n), dtype=torch.uint8, device=self.device)
        trunc = torch.empty((T, n), dtype=torch.uint8, device=self.device)
        _capi.check(self._L.ta_rollout(self._h, _ptr(actions), _ACT_DTYPES[actions.dtype], T, _ptr(obs), _ptr(rew),
                                       _ptr(term), _ptr(trunc), self._stream()), "ta_rollout")
        return obs, rew, term.view(torch.bool), trunc.view(torch.bool)

    def reset_obs_template(self) -> torch.Tensor:
        """The observation reset() always produces (deterministic _gen_grid): [V,V,3] uint8.
        Callers that want gymnasium's same-step autoreset semantics select it where done."""
        if self._reset_template is None:
            probe = TwoarmyVecEnv(self.version, 1, self.view, self.device)
            self._reset_template = probe.reset()[0].clone()
            probe.close()
        return self._reset_template

    # ------------------------------------------------------------------ featurise (Env_transact)
    def state_matrix(self, want_codes: bool = False):
        """Env_transact.matrix_env / data_env (soa/env_buffer.py:300-334) for every env:
        (matrix float32 [N,289], place float32 [N,2] = (y,x)[, codes uint8 [N,289]])."""
        n = self.num_envs
        mat = torch.empty((n, 289), dtype=torch.float32, device=self.device)
        place = torch.empty((n, 2), dtype=torch.float32, device=self.device)
        codes = torch.empty((n, 289), dtype=torch.uint8, device=self.device) if want_codes else None
        _capi.check(self._L.ta_state_matrix(self._h, _ptr(codes), _ptr(mat), _ptr(place), self._stream()),
                    "ta_state_matrix")
        return (mat, place, codes) if want_codes else (mat, place)

    def stack_roll(self, s_stack: torch.Tensor, p_stack: torch.Tensor, init_mask: Optional[torch.Tensor] = None,
                   init: bool = False):
        """In-place frame-stack update of s_stack [N,5,289] / p_stack [N,5,2]
        (train_ppo.py:116-121; init=True tiles the current frame, env_buffer.py:420-423)."""
        assert s_stack.is_contiguous() and p_stack.is_contiguous()
        assert s_stack.shape == (self.num_envs, 5, 289) and p_stack.shape == (self.num_envs, 5, 2)
        assert s_stack.dtype == torch.float32 and p_stack.dtype == torch.float32
        if init_mask is not None:
            init_mask = init_mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _capi.check(self._L.ta_stack_roll(self._h, _ptr(s_stack), _ptr(p_stack), _ptr(init_mask), int(init),
                                          self._stream()), "ta_stack_roll")

    def stack_roll_codes(self, s_codes: torch.Tensor, p_stack: torch.Tensor, init_mask: Optional[torch.Tensor] = None,
                         init: bool = False):
        """stack_roll on uint8 codes [N,5,289] (what the device rollout buffer stores)."""
        assert s_codes.is_contiguous() and p_stack.is_contiguous()
        assert s_codes.shape == (self.num_envs, 5, 289) and p_stack.shape == (self.num_envs, 5, 2)
        assert s_codes.dtype == torch.uint8 and p_stack.dtype == torch.float32
        if init_mask is not None:
            init_mask = init_mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _capi.check(self._L.ta_stack_roll_codes(self._h, _ptr(s_codes), _ptr(p_stack), _ptr(init_mask), int(init),
                                                self._stream()), "ta_stack_roll_codes")

    def stack_push(self, s_prev: Optional[torch.Tensor], s_out: torch.Tensor, p_prev: Optional[torch.Tensor],
                   p_out: Optional[torch.Tensor], prev_done: Optional[torch.Tensor] = None, init_all: bool = False):
        """Out-of-place frame-stack push (ta_stack_push): s_out = roll(base) + current frame with
        base = tiled reset frame where prev_done / init_all, else s_prev.  uint8 codes or float32."""
        assert s_out.is_contiguous() and s_out.shape == (self.num_envs, 5, 289) and s_out.dtype in (torch.uint8, torch.float32)
        if s_prev is not None:
            assert s_prev.is_contiguous() and s_prev.shape == s_out.shape and s_prev.dtype == s_out.dtype
        if p_out is not None:
            assert p_out.is_contiguous() and p_out.shape == (self.num_envs, 5, 2) and p_out.dtype == torch.float32
        if p_prev is not None:
            assert p_prev.is_contiguous() and p_prev.dtype == torch.float32
        if prev_done is not None:
            prev_done = prev_done.to(device=self.device, dtype=torch.uint8).contiguous()
        _capi.check(self._L.ta_stack_push(self._h, _ptr(s_prev), _ptr(s_out), _ptr(p_prev), _ptr(p_out), _ptr(prev_done),
                                          int(init_all), 1 if s_out.dtype == torch.uint8 else 0, self._stream()), "ta_stack_push")

    def render(self, env_ids: Optional[torch.Tensor] = None, tile_size: int = 32, highlight: bool = False) -> torch.Tensor:
        """get_full_render() of the selected envs: uint8 [M, 17*tile_size, 17*tile_size, 3]."""
        from . import render as _render
        return _render.render(self, env_ids, tile_size, highlight)

    # ------------------------------------------------------------------ state access
    def export_state(self) -> np.ndarray:
        buf = torch.empty(self.num_envs * STATE_DTYPE.itemsize, dtype=torch.uint8, device=self.device)
        _capi.check(self._L.ta_export_state(self._h, _ptr(buf), self._stream()), "ta_export_state")
        return buf.cpu().numpy().view(STATE_DTYPE).copy()

    def import_state(self, state: np.ndarray):
        assert state.dtype == STATE_DTYPE and state.shape == (self.num_envs,)
        buf = torch.from_numpy(np.ascontiguousarray(state).view(np.uint8).copy()).to(self.device)
        _capi.check(self._L.ta_import_state(self._h, _ptr(buf), self._stream()), "ta_import_state")
        torch.cuda.current_stream(self.device).synchronize()
