from ..core import Wrapper


class TimeLimit(Wrapper):
    def __init__(self, env, max_episode_steps=None):
        super().__init__(env)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = None

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


class RecordVideo(Wrapper):
    def __init__(self, env, *a, **k):
        super().__init__(env)
