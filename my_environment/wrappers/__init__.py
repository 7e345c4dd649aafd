"""`from my_environment.wrappers import EpisodeAnalyzer, RewardAnnealing` (reference: my_environment/wrappers/__init__.py)."""
from my_environment.wrappers.wrappers import *  # noqa: F401,F403
from my_environment.wrappers import wrappers  # noqa: F401
