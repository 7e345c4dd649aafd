"""rl_rocket_b200 -- B200-native batched implementation of the RL_rocket environment step.

Importing the package is cheap (no CUDA work).  The environment classes need the in-tree CUDA extension
(rl_rocket_b200/_C/librlrocket.so, `python -m rl_rocket_b200.csrc.build`) and a GPU: there is no CPU fallback.
Gym ids `my_environment/Falcon3DOF-v0` and `my_environment/Falcon6DOF-v0` (reference: my_environment/__init__.py:4-12)
are registered when gym is importable.
"""
from .config import EnvSpec  # noqa: F401

__all__ = ["EnvSpec", "Rocket", "Rocket6DOF", "Rocket1DOF", "RocketVecEnv", "RocketBatch"]


def __getattr__(name):
    if name in ("Rocket", "Rocket6DOF", "Rocket1DOF"):
        from . import envs
        return getattr(envs, name)
    if name == "RocketVecEnv":
        from .vec_env import RocketVecEnv
        return RocketVecEnv
    if name == "RocketBatch":
        from .engine import RocketBatch
        return RocketBatch
    raise AttributeError(name)


def _register_gym_ids(prefix="rl_rocket_b200.envs"):
    try:
        from gym.envs.registration import register
    except Exception:  # noqa: BLE001 - gym absent: nothing to register
        return False
    for gid, cls in (("my_environment/Falcon3DOF-v0", "Rocket"), ("my_environment/Falcon6DOF-v0", "Rocket6DOF")):
        try:
            register(id=gid, entry_point=f"{prefix}:{cls}")
        except Exception:  # noqa: BLE001 - already registered
            pass
    return True


_register_gym_ids()
