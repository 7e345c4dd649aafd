"""`from my_environment.envs import Rocket, Rocket6DOF` (reference: my_environment/envs/__init__.py, rocket_env.py:19, :505)."""
from rl_rocket_b200.envs import Rocket, Rocket1DOF, Rocket6DOF  # noqa: F401
from my_environment.envs import rocket_env  # noqa: F401,E402

__all__ = ["Rocket", "Rocket6DOF", "Rocket1DOF"]
