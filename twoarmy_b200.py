"""Import alias: `import twoarmy_b200` loads the package whose directory name
(goal-conditioned-reinforcement-learning-with-environmental-and-policy-priors_b200) is not a
valid identifier, and replaces this module with it."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module(
    "goal-conditioned-reinforcement-learning-with-environmental-and-policy-priors_b200")
sys.modules[__name__] = _pkg
