"""Wrappers of the reference (my_environment/wrappers/wrappers.py) that sit on the step path, for the single-env facades.
The vectorised equivalents are constructor flags of RocketVecEnv (discrete_actions=..., reward_annealing=...), where the
action-table lookup and the reward switch are fused into the step kernel.

  DiscreteActions3DOF  wrappers.py:24-48     RewardAnnealing  wrappers.py:68-86     GaudetStateObs  wrappers.py:51-66
  OneDOF  README.md:4-8 (spec only)
Logging / video wrappers (RecordVideoFigure, EpisodeAnalyzer*, wrappers.py:88-349) are out of scope of the hot path;
they exist here as inert pass-through classes so that the reference's training scripts import and run unmodified
(main_6DOF.py:15, :41-46; main.py:20, :101-103).
"""
import numpy as np

from . import spaces
from .envs import Rocket, Rocket6DOF


class _Wrapper(spaces.GymEnv):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, **kw):
        return self.env.reset(**kw)

    def step(self, action):
        return self.env.step(action)

    def seed(self, seed=None):
        return self.env.seed(seed)

    def close(self):
        return self.env.close()


class DiscreteActions3DOF(_Wrapper):
    def __init__(self, env, disc_to_cont=[[0, -1], [-1, +1], [0, +1], [+1, +1]]):
        super().__init__(env)
        self.disc_to_cont = disc_to_cont
        self.action_space = spaces.Discrete(len(disc_to_cont))
        self.observation_space = env.observation_space

    def action(self, act):
        return np.asarray(self.disc_to_cont[act])

    def step(self, act):
        return self.env.step(self.action(act))


class RewardAnnealing(_Wrapper):
    def __init__(self, env, thrust_penalty: float = 0.01):
        super().__init__(env)
        self.xi = self.env.reward_coefficients.get("xi", thrust_penalty)
        self.action_space, self.observation_space = env.action_space, env.observation_space

    def step(self, action):
        obs, __, done, info = self.env.step(action)
        old = info["rewards_dict"]
        rewards_dict = {key: old[key] for key in ["attitude_constraint", "rew_goal"]}
        thr = action[2] + 1 if isinstance(self.env.unwrapped, Rocket6DOF) else action[1] + 1
        rewards_dict["thrust_penalty"] = -self.xi * thr
        reward = sum(rewards_dict.values())
        info["rewards_dict"] = rewards_dict
        return obs, reward, done, info


class GaudetStateObs(_Wrapper):
    """wrappers.py:51-66 with its two defects repaired (it calls `compute_vtarg`, which does not exist -- the method is
    `_compute_vtarg`, rocket_env.py:219 -- and hands it the NORMALISED observation although the method works on the physical
    state): observation = float32([vx - vt_x, vz - vt_z, t_go, z]) of the physical float32 state, computed in the step
    kernel's epilogue (RLR_OBS_GAUDET), pinned by tests/golden/traj_3dof_gaudet.npz."""

    def __init__(self, env):
        super().__init__(env)
        base = env.unwrapped
        if not isinstance(base, Rocket) or base._dof != 3:
            raise TypeError("GaudetStateObs wraps the 3DOF Rocket environment (wrappers.py:52)")
        base._set_obs_mode("gaudet")
        self.observation_space = spaces.Box(low=-1, high=1, shape=(4,))
        self.action_space = env.action_space

    def observation(self, observation):
        return observation                                         # already the 4-vector: the kernel wrote it


class _Inert(_Wrapper):
    """Logging / video wrappers of the reference (wrappers.py:88-349): pass-through here (rendering, wandb and video capture
    are outside the hot path); every constructor argument of the reference is accepted and ignored."""

    def __init__(self, env, *args, **kwargs):
        super().__init__(env)
        self.action_space, self.observation_space = env.action_space, env.observation_space
        self.episode_id = 0

    def reset(self, **kw):
        self.episode_id += 1
        return self.env.reset(**kw)


class EpisodeAnalyzer(_Inert):
    pass


class EpisodeAnalyzer6DOF(_Inert):
    pass


class RecordVideoFigure(_Inert):
    pass


class TimeLimit(_Wrapper):
    """gym.wrappers.TimeLimit semantics (main_6DOF.py:21, main.py:65) for use when gym itself is not installed."""

    def __init__(self, env, max_episode_steps):
        super().__init__(env)
        self._max_episode_steps, self._elapsed_steps = max_episode_steps, None
        self.action_space, self.observation_space = env.action_space, env.observation_space

    def step(self, action):
        o, r, d, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            info["TimeLimit.truncated"] = not d
            d = True
        return o, r, d, info

    def reset(self, **kw):
        self._elapsed_steps = 0
        return self.env.reset(**kw)
