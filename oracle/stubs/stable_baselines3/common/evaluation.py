import numpy as np


def evaluate_policy(model, env, n_eval_episodes=10, deterministic=True, render=False, **kw):
    env = model._wrap_env(env)
    rets, cur, obs = [], np.zeros(env.num_envs), env.reset()
    while len(rets) < n_eval_episodes:
        a, _ = model.predict(obs, deterministic=deterministic)
        obs, r, d, infos = env.step(a)
        cur += r
        for i in range(env.num_envs):
            if d[i]:
                rets.append(cur[i])
                cur[i] = 0
    return float(np.mean(rets)), float(np.std(rets))
