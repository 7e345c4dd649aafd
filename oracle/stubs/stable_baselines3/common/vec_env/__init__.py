from .base_vec_env import VecEnv  # noqa: F401
from .dummy_vec_env import DummyVecEnv  # noqa: F401
