import numpy as np

from .base_vec_env import VecEnv


class DummyVecEnv(VecEnv):
    """dummy_vec_env.py: sequential envs; on done the terminal observation goes to info and the env is reset at once."""

    def __init__(self, env_fns):
        self.envs = [fn() for fn in env_fns]
        e = self.envs[0]
        super().__init__(len(self.envs), e.observation_space, e.action_space)
        self.actions = None

    def seed(self, seed=None):
        return [e.seed(None if seed is None else seed + i) for i, e in enumerate(self.envs)]

    def reset(self):
        return np.stack([np.asarray(e.reset()) for e in self.envs])

    def step_async(self, actions):
        self.actions = actions

    def step_wait(self):
        obs, rews, dones, infos = [], [], [], []
        for e, a in zip(self.envs, self.actions):
            o, r, d, info = e.step(a)
            if d:
                info["terminal_observation"] = o
                o = e.reset()
            obs.append(np.asarray(o)), rews.append(r), dones.append(d), infos.append(info)
        return np.stack(obs), np.asarray(rews, dtype=np.float32), np.asarray(dones, dtype=bool), infos

    def close(self):
        for e in self.envs:
            e.close()

    def get_attr(self, name, indices=None):
        return [getattr(e, name) for e in self.envs]

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self.envs]
