import numpy as np

from ..common.monitor import Monitor
from ..common.vec_env.dummy_vec_env import DummyVecEnv
from ..common.vec_env.base_vec_env import VecEnv


class PPO:
    """on-policy algorithm skeleton: _wrap_env (Monitor + DummyVecEnv, base_class.py), set_env's space checks, learn() =
    collect_rollouts with a random policy (n_steps per rollout, on_policy_algorithm.py)."""

    def __init__(self, policy, env, n_steps=64, seed=None, verbose=0, **kw):
        self.policy_type, self.n_steps, self.seed, self.verbose = policy, n_steps, seed, verbose
        self.env = self._wrap_env(env)
        self.observation_space, self.action_space = self.env.observation_space, self.env.action_space
        self.n_envs = self.env.num_envs
        if seed is not None:
            self.action_space.seed(seed)
            self.env.seed(seed)
        self._last_obs = None
        self.num_timesteps = 0
        self.ep_info_buffer = []
        self.rollout = None

    @staticmethod
    def _wrap_env(env):
        if not isinstance(env, VecEnv) and not (hasattr(env, "step_async") and hasattr(env, "num_envs")):
            if not isinstance(env, Monitor):
                env = Monitor(env)
            env = DummyVecEnv([lambda: env])
        return env

    def get_env(self):
        return self.env

    def set_env(self, env):
        env = self._wrap_env(env)
        assert env.observation_space.shape == self.observation_space.shape, "observation spaces do not match"
        assert type(env.action_space) is type(self.action_space) and env.action_space.shape == self.action_space.shape, \
            "action spaces do not match"
        assert env.num_envs == self.n_envs
        self.env, self._last_obs = env, None

    def predict(self, obs, state=None, episode_start=None, deterministic=False):
        n = np.asarray(obs).shape[0] if np.asarray(obs).ndim > len(self.observation_space.shape) else 1
        a = np.stack([self.action_space.sample() for _ in range(n)])
        return a, None

    def learn(self, total_timesteps, callback=None, **kw):
        if self._last_obs is None:
            self._last_obs = self.env.reset()
        obs_l, rew_l, done_l = [], [], []
        t = 0
        while t < total_timesteps:
            for _ in range(self.n_steps):
                a, _s = self.predict(self._last_obs)
                obs, rew, done, infos = self.env.step(a)
                assert np.asarray(obs).shape == (self.n_envs,) + tuple(self.observation_space.shape)
                for i, d in enumerate(done):
                    if d:
                        assert "terminal_observation" in infos[i]
                    if "episode" in infos[i]:
                        self.ep_info_buffer.append(infos[i]["episode"])
                obs_l.append(np.asarray(obs)), rew_l.append(np.asarray(rew)), done_l.append(np.asarray(done))
                self._last_obs = obs
                t += self.n_envs
                self.num_timesteps += self.n_envs
        self.rollout = {"obs": np.stack(obs_l), "rewards": np.stack(rew_l), "dones": np.stack(done_l)}
        return self

    def save(self, path):
        pass
