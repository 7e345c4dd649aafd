"""`my_environment.wrappers.wrappers` (reference: wrappers.py:1-8 `__all__`): the step-path wrappers are the ones of
rl_rocket_b200.wrappers; the logging / video wrappers are inert pass-throughs (outside the hot path)."""
from rl_rocket_b200.wrappers import (DiscreteActions3DOF, EpisodeAnalyzer, EpisodeAnalyzer6DOF, GaudetStateObs,  # noqa: F401
                                     RecordVideoFigure, RewardAnnealing)

__all__ = ["DiscreteActions3DOF", "GaudetStateObs", "RewardAnnealing", "RecordVideoFigure", "EpisodeAnalyzer6DOF", "EpisodeAnalyzer"]
