"""`my_environment.envs.rocket_env` (reference: rocket_env.py): the two environment classes, backed by the CUDA step path."""
from rl_rocket_b200.envs import Rocket, Rocket1DOF, Rocket6DOF  # noqa: F401

__all__ = ["Rocket", "Rocket6DOF"]
