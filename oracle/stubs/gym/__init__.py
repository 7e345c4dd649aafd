"""Minimal stand-in for gym 0.21 -- TEST INFRASTRUCTURE ONLY.

gym is not installable in this image (no network).  The live reference
(/root/reference/my_environment) imports `gym.Env`, `gym.spaces.Box`,
`gym.envs.registration.register` and `gym.wrappers.TimeLimit` at module scope;
this stub provides just those so the reference sources run UNMODIFIED.
Semantics follow gym 0.21 as recalled in SURVEY.md Appendix D.
"""
from . import spaces  # noqa: F401
from .core import Env, Wrapper, ActionWrapper, ObservationWrapper, RewardWrapper  # noqa: F401
from .envs.registration import register, make  # noqa: F401
from . import logger  # noqa: F401
