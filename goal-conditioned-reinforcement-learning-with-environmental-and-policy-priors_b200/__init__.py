"""B200-native batched Twoarmy hot path (step + gen_obs -> rollout buffer -> advantages).

The directory name is the one the build contract prescribes and is not a valid Python
identifier; import it as `twoarmy_b200` (the shim at the repo root) or with
importlib.import_module("goal-conditioned-reinforcement-learning-with-environmental-and-policy-priors_b200").
"""
from . import _capi  # noqa: F401
from ._capi import TwoarmyLibraryError, build, launch_count  # noqa: F401
from .vec_env import ENV_IDS, STATE_DTYPE, TwoarmyVecEnv  # noqa: F401

__all__ = ["TwoarmyVecEnv", "ENV_IDS", "STATE_DTYPE", "TwoarmyLibraryError", "build", "launch_count"]
