"""Minimal stand-in for stable-baselines3 1.6 -- TEST INFRASTRUCTURE ONLY (SB3 is not installable in this image).

Just enough of the public surface the reference's scripts touch (main.py:10-17, main_6DOF.py:10-13,
test_6DOF_sb_integration.py:5-8) to execute their environment-construction and rollout-collection code paths against the
B200-backed environments: Monitor, DummyVecEnv, check_env, evaluate_policy and an "algorithm" whose policy is
`action_space.sample()` (no learning -- the policy network is outside the hot path).  Semantics follow SB3 1.6 as recalled in
SURVEY.md Appendix D.  The product package never imports this.
"""
from .ppo import PPO  # noqa: F401
from .ppo import PPO as A2C, PPO as DQN, PPO as TD3  # noqa: F401
