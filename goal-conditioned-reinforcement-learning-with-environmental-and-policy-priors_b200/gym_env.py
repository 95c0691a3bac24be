"""Single-env gym facade: the drop-in for `gym.make("MiniGrid-twoarmy-17x17-v4"|"-v6", **kw)`.

Mirrors the reference boundary (gym_minigrid/__init__.py:10-20, Twoarmy_v4.__init__
twoarmy_v4.py:9-36, MiniGridEnv minigrid.py:835-1590): same constructor kwargs, reset()/step()
signatures and return types, and the attributes the reference's own callers reach through the
boundary: env.grid.grid[i].type / .height, env.agent_pos, env.goal_pos,
env.obstacles[k].cur_pos, env.actions.*, env.max_steps, env.step_count, env.mission,
env.get_full_render().  The transition itself runs in the CUDA library on a batch of one.

RNG.  The reference draws from the GLOBAL legacy numpy stream (np.random.choice at
twoarmy_v4.py:117,149,184,190,215,221,303,310).  rng="numpy" (default) keeps that contract
bit for bit: the step is executed in the library's verification mode, and each call site
the kernel reports as executed is fed the value np.random.choice returns, in program order --
so `np.random.seed(s)` followed by the same actions gives the reference's trajectory.
rng="philox" uses the library's counter-based production stream instead.
"""
from __future__ import annotations

from enum import IntEnum
from typing import Optional

import numpy as np
import torch

from . import _capi
from .vec_env import ENV_IDS, GOAL_POS, MAX_STEPS, MISSION, TwoarmyVecEnv

try:  # the facade subclasses gym.Env when a gym is installed, so isinstance checks hold
    import gym as _gym  # type: ignore
    _EnvBase = _gym.Env
except Exception:  # pragma: no cover - gym is not part of this image
    try:
        import gymnasium as _gym  # type: ignore
        _EnvBase = _gym.Env
    except Exception:
        _gym = None
        _EnvBase = object

OBJECT_TO_IDX = {"unseen": 0, "empty": 1, "wall": 2, "floor": 3, "door": 4, "key": 5, "ball": 6, "box": 7,
                 "goal": 8, "lava": 9, "agent": 10, "subgoal": 11}   # minigrid.py:45-58
COLOR_TO_IDX = {"red": 0, "green": 1, "blue": 2, "purple": 3, "yellow": 4, "grey": 5}  # minigrid.py:40
_CELL = {1: ("wall", "grey"), 2: ("ball", "yellow"), 3: ("goal", "green")}
# slot -> (lo, hi) of np.random.choice(range(lo, hi), 1)
_REWARDS = (-0.01, -0.1, -0.9, 0.2, 0.9)  # twoarmy_v4.py:180,229,240,284,295
_REWARD_BY_F32 = {np.float32(r).item(): r for r in _REWARDS}
_SLOT_RANGE = {0: (0, 10), 1: (9, 13), 2: (6, 10), 3: (6, 10), 4: (4, 5), 5: (0, 2), 6: (0, 2)}


class Actions(IntEnum):  # MiniGridEnv.Actions, minigrid.py:849-864
    left = 0
    right = 1
    up = 2
    down = 3
    drop = 4
    toggle = 5
    done = 6


class _Space:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class WorldObjView:
    """What callers read off a grid cell / ball object (env_buffer.py:302-330)."""

    def __init__(self, type_, color, cur_pos=None):
        self.type, self.color, self.cur_pos, self.init_pos = type_, color, cur_pos, cur_pos

    def encode(self):
        return (OBJECT_TO_IDX[self.type], COLOR_TO_IDX[self.color], 0)

    def can_overlap(self):
        return self.type == "goal"


class GridView:
    """Read-only snapshot with Grid's public surface (minigrid.py:554-772)."""

    def __init__(self, codes: np.ndarray):
        self.width = self.height = 17
        self.grid = [None if c == 0 else WorldObjView(*_CELL[int(c)]) for c in codes]

    def get(self, i, j):
        assert 0 <= i < self.width and 0 <= j < self.height
        return self.grid[j * self.width + i]

    def encode(self, vis_mask=None):
        arr = np.zeros((self.width, self.height, 3), dtype="uint8")
        for i in range(self.width):
            for j in range(self.height):
                if vis_mask is None or vis_mask[i, j]:
                    v = self.get(i, j)
                    arr[i, j] = (1, 0, 0) if v is None else v.encode()
        return arr


class TwoarmyEnv(_EnvBase):
    metadata = {"render_modes": ["human", "rgb_array", "single_rgb_array"], "render_fps": 10}
    version = 4

    def __init__(self, size=17, agent_pos=(3, 15), goal_pos=(14, 2), agent_view_size: int = 17,
                 render_mode: Optional[str] = None, highlight: bool = False, tile_size: int = 32,
                 agent_pov: bool = False, device="cuda:0", rng: str = "numpy", philox_seed: int = 9981, **kwargs):
        if size != 17 or tuple(agent_pos) != (3, 15) or tuple(goal_pos) != (14, 2):
            raise NotImplementedError("only the registered 17x17 layout with the default agent/goal "
                                      "positions is built (gym_minigrid/__init__.py:10-20)")
        assert rng in ("numpy", "philox")
        self.render_mode, self.highlight, self.tile_size, self.agent_pov = render_mode, highlight, tile_size, agent_pov
        self.actions = Actions
        self.action_space = _Space(n=len(Actions)) if _gym is None else _gym.spaces.Discrete(len(Actions))
        self.agent_view_size = agent_view_size
        self.observation_space = _Space(shape=(agent_view_size, agent_view_size, 3), dtype="uint8")
        self.reward_range = (0, 1)
        self.width = self.height = size
        self.max_steps = MAX_STEPS
        self.see_through_walls = True
        self.mission = MISSION
        self.goal_pos = GOAL_POS
        self.agent_dir = 3
        self.carrying = None
        self.n_obstacles = 3
        self._rng = rng
        self._vec = TwoarmyVecEnv(self.version, 1, agent_view_size, device=device, seed=philox_seed, autoreset=False)
        self._state = None
        self.reset()  # MiniGridEnv.__init__ ends with self.reset() (minigrid.py:945)

    # ---- state mirror ------------------------------------------------------------------------
    def _refresh(self):
        self._state = self._vec.export_state()[0]

    def _flag(self, bit):
        return bool(self._state["flags"] & bit)

    @property
    def grid(self):
        return GridView(self._state["grid"])

    @property
    def agent_pos(self):
        return (int(self._state["agent_x"]), int(self._state["agent_y"]))

    @property
    def step_count(self):
        return int(self._state["step_count"])

    @property
    def step_move(self):
        return int(self._state["step_move"])

    @property
    def risk_count(self):
        return int(self._state["risk_count"])

    pone = property(lambda self: self._flag(_capi.F_PONE))
    patrol = property(lambda self: self._flag(_capi.F_PATROL))
    up1 = property(lambda self: self._flag(_capi.F_UP1))
    right2 = property(lambda self: self._flag(_capi.F_RIGHT2))
    Update_horizontal = property(lambda self: self._flag(_capi.F_UPD_H))
    Update_longitudinal = property(lambda self: self._flag(_capi.F_UPD_L))
    first_to_room2 = property(lambda self: self._flag(_capi.F_FIRST_ROOM2))

    def _balls(self, lo, hi):
        out = []
        for k in range(lo, hi):
            x, y = (int(v) for v in self._state["balls"][k])
            out.append(WorldObjView("ball", "yellow", None if x == 255 else (x, y)))
        return out

    obstacles = property(lambda self: self._balls(0, 3))
    obstacles1 = property(lambda self: self._balls(3, 6))
    obstacles2 = property(lambda self: self._balls(6, 10))

    # ---- gym API -----------------------------------------------------------------------------
    def _obs_dict(self, image: np.ndarray):
        return {"image": image, "direction": self.agent_dir, "mission": self.mission}  # minigrid.py:1494

    def reset(self, *, seed=None, return_info=False, options=None):
        """minigrid.py:947-980: regenerates grid, agent and step_count; the Twoarmy flags carry
        over exactly as in the reference.  `seed` is accepted and ignored like the reference's
        gym seeding (no draw is made from it with the default agent/goal positions)."""
        img = self._vec.reset()[0].cpu().numpy()
        self._refresh()
        obs = self._obs_dict(img)
        return obs if not return_info else (obs, {})

    def gen_obs(self):
        return self._obs_dict(self._vec.observe()[0].cpu().numpy())

    def step(self, action):
        a = int(action)
        if a >= len(Actions):  # twoarmy_v4.py:84-85
            a = 0
        if a not in (0, 1, 2, 3, 6):
            if a in (4, 5):  # minigrid.py:1397: Actions has no `forward`
                raise AttributeError("forward")
            raise ValueError(f"Unknown action: {action}")
        act = torch.tensor([a], dtype=torch.int32)
        if self._rng == "philox":
            obs, rew, te, tr, _ = self._vec.step(act)
        else:
            # verification mode with lazily discovered call sites: re-run the step from the saved
            # state until every site the kernel executed has been given its np.random value
            saved = self._vec.export_state()
            draws = np.full((1, 8), 0xFF, np.uint8)
            known = 0
            while True:
                obs, rew, te, tr, info = self._vec.step(act, draws=torch.from_numpy(draws))
                consumed = int(info["consumed"][0])
                missing = consumed & ~known
                if not missing:
                    break
                slot = (missing & -missing).bit_length() - 1  # first executed site without a value
                lo, hi = _SLOT_RANGE[slot]
                draws[0, slot] = np.random.choice(range(lo, hi), 1).item()
                known |= 1 << slot
                self._vec.import_state(saved)
        self._refresh()
        if self._state["error"]:
            raise AssertionError(f"env error bits {int(self._state['error'])} (see TA_ENV_ERR_*)")
        reward = _REWARD_BY_F32[float(rew[0])]  # the reference returns these Python floats
        return self._obs_dict(obs[0].cpu().numpy()), reward, bool(te[0]), bool(tr[0]), {}

    # ---- rendering (host side; not part of the accelerated path) -----------------------------
    def get_full_render(self, highlight=None, tile_size=None):
        """minigrid.py:1514-1563: the RGB picture of the whole grid, (17*tile, 17*tile, 3) uint8,
        pixel-identical to the reference's software rasteriser (ta_render blits the cached tiles)."""
        from . import render as _render
        ts = int(tile_size or self.tile_size)
        hl = self.highlight if highlight is None else highlight
        return _render.render(self._vec, None, ts, bool(hl))[0].cpu().numpy()

    def render(self, *a, **k):
        return self.get_full_render()

    def close(self):
        self._vec.close()


class Twoarmy_v4(TwoarmyEnv):  # gym_minigrid/envs/twoarmy_v4.py
    version = 4


class Twoarmy_v6(TwoarmyEnv):  # gym_minigrid/envs/twoarmy_v6.py
    version = 6


_ENTRY = {"MiniGrid-twoarmy-17x17-v4": Twoarmy_v4, "MiniGrid-twoarmy-17x17-v6": Twoarmy_v6}


def make(env_id: str, **kwargs):
    """gym.make for the two registered ids; kwargs={"size": 17} as in gym_minigrid/__init__.py."""
    if env_id not in _ENTRY:
        raise KeyError(f"unknown env id {env_id!r}; registered: {sorted(_ENTRY)}")
    args = {"size": 17}
    args.update(kwargs)
    return _ENTRY[env_id](**args)


def register_minigrid_envs():
    """Registers the two ids with gym / gymnasium when one is installed (same ids and kwargs as
    gym_minigrid/__init__.py:6-20); otherwise `make` above is the entry point."""
    if _gym is None:
        return False
    from gym.envs.registration import register  # type: ignore
    pkg = __name__
    for env_id, cls in _ENTRY.items():
        register(id=env_id, entry_point=f"{pkg}:{cls.__name__}", kwargs={"size": 17})
    return True


assert set(_ENTRY) == set(ENV_IDS)
