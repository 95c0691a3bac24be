"""Test-only import shim for the Python reference (used ONLY by make_golden.py, and only
where /root/reference exists).  The reference needs `gym`, `matplotlib` and `turtle`, none
of which are installed; these stubs carry no arithmetic -- every number on the path still
comes from the reference's own files and numpy."""
import importlib
import sys
import types
import typing

import numpy as np

REFERENCE_ROOT = "/root/reference"


def find_root():
    """Where the Python reference lives: /root/reference in the build container; on the GPU box the
    git-ignored copy __graft_entry__.build() ships under baseline/_ref (used by bench.py's CPU arm only)."""
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    shipped = os.path.join(os.path.dirname(os.path.dirname(here)), "baseline", "_ref")
    for root in (os.environ.get("TA_REFERENCE_ROOT"), REFERENCE_ROOT, shipped):
        if root and os.path.isfile(os.path.join(root, "gym_minigrid", "minigrid.py")):
            return root
    return None


def install(root=None):
    if "gym_minigrid" in sys.modules:
        return
    root = root or find_root()
    if root is None:
        raise FileNotFoundError("the Python reference is neither at /root/reference nor shipped under baseline/_ref")
    gym = types.ModuleType("gym")

    class Env:
        def reset(self, *, seed=None, **kw):
            return None

    gym.Env = Env
    spaces = types.ModuleType("gym.spaces")
    T = typing.TypeVar("T")

    class Space(typing.Generic[T]):
        def __init__(self, shape=None, dtype=None, seed=None):
            self.shape, self.dtype = shape, dtype

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    class Dict(Space):
        def __init__(self, d):
            self.spaces = d

    class Discrete(Space):
        def __init__(self, n):
            self.n = n

    spaces.Space, spaces.Box, spaces.Dict, spaces.Discrete = Space, Box, Dict, Discrete
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = lambda seed=None: (np.random.default_rng(seed), seed)
    seeding.RandomNumberGenerator = np.random.Generator
    utils.seeding = seeding
    envs = types.ModuleType("gym.envs")
    reg = types.ModuleType("gym.envs.registration")
    registry = {}

    def register(id, entry_point, kwargs=None, **kw):
        registry[id] = (entry_point, kwargs or {})

    def make(id, **kw):
        ep, base = registry[id]
        mod, cls = ep.split(":")
        args = dict(base)
        args.update(kw)
        return getattr(importlib.import_module(mod), cls)(**args)

    reg.register = register
    envs.registration = reg
    core = types.ModuleType("gym.core")

    class Wrapper:
        pass

    core.Wrapper = core.ObservationWrapper = Wrapper
    gym.Wrapper = gym.ObservationWrapper = Wrapper
    gym.spaces, gym.utils, gym.envs, gym.core, gym.make = spaces, utils, envs, core, make
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    turtle = types.ModuleType("turtle")
    turtle.right = None
    for name, mod in [("gym", gym), ("gym.spaces", spaces), ("gym.utils", utils),
                      ("gym.utils.seeding", seeding), ("gym.envs", envs),
                      ("gym.envs.registration", reg), ("gym.core", core),
                      ("matplotlib", mpl), ("matplotlib.pyplot", plt), ("turtle", turtle)]:
        sys.modules[name] = mod
    sys.path.insert(0, root)
    sys.path.insert(0, root + "/soa")
    import gym_minigrid

    gym_minigrid.register_minigrid_envs()
