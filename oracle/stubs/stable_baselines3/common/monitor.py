import time

import gym


class Monitor(gym.Wrapper):
    """monitor.py: per-episode return / length / time put into info["episode"] when the episode ends."""

    def __init__(self, env, filename=None, allow_early_resets=True, reset_keywords=(), info_keywords=()):
        super().__init__(env)
        self.t_start = time.time()
        self.rewards, self.needs_reset = None, True
        self.episode_returns, self.episode_lengths, self.episode_times = [], [], []
        self.total_steps = 0

    def reset(self, **kw):
        self.rewards, self.needs_reset = [], False
        return self.env.reset(**kw)

    def step(self, action):
        if self.needs_reset:
            raise RuntimeError("Tried to step environment that needs reset")
        obs, reward, done, info = self.env.step(action)
        self.rewards.append(reward)
        if done:
            self.needs_reset = True
            ep = {"r": round(sum(self.rewards), 6), "l": len(self.rewards), "t": round(time.time() - self.t_start, 6)}
            self.episode_returns.append(ep["r"]), self.episode_lengths.append(ep["l"]), self.episode_times.append(ep["t"])
            info["episode"] = ep
        self.total_steps += 1
        return obs, reward, done, info

    def get_episode_rewards(self):
        return self.episode_returns

    def get_episode_lengths(self):
        return self.episode_lengths
