"""Executed in a subprocess by tests/test_gpu_dropin.py with `oracle/stubs` (gym 0.21 / stable-baselines3 1.6 / wandb stand-ins)
and the repo root on PYTHONPATH.  It walks the environment-construction and rollout code paths of the reference's training and
test scripts against the `my_environment` alias package -- the function bodies below restate, line for line in meaning,
    main_6DOF.py:18-46, :52-105     make_env / make_annealed_env / make_eval_env / PPO(...) / learn / set_env(annealed) / learn
    main.py:48-71, :101-105, :130-175   3DOF make_env (+ Monitor kwargs), eval env in DummyVecEnv, RewardAnnealing phase, set_env
    test_6DOF_sb_integration.py:13-36   check_env, PPO('MlpPolicy', env), evaluate_policy
    test_6DOF_environment.py:13-31      fixed-action episode until done
(the scripts themselves live in /root/reference, which does not exist on the GPU box).  Prints one JSON line of facts the
test asserts on."""
import json

import numpy as np

import my_environment  # noqa: F401  (registers the gym ids, my_environment/__init__.py)
import gym
import wandb
from gym.wrappers import TimeLimit, RecordVideo
from stable_baselines3 import PPO
from stable_baselines3.common.callbacks import EvalCallback
from stable_baselines3.common.env_checker import check_env
from stable_baselines3.common.evaluation import evaluate_policy
from stable_baselines3.common.monitor import Monitor
from stable_baselines3.common.vec_env.dummy_vec_env import DummyVecEnv
from wandb.integration.sb3 import WandbCallback

from my_environment.envs import Rocket6DOF
from my_environment.wrappers import EpisodeAnalyzer, RewardAnnealing
from my_environment.wrappers.wrappers import DiscreteActions3DOF, GaudetStateObs, RecordVideoFigure

from rl_rocket_b200.config import env_config_3dof, env_config_6dof

env_config = env_config_6dof()                                             # configuration_file.py:4-34
sb3_config = {"policy_type": "MlpPolicy", "total_timesteps": 192, "max_ep_timesteps": 800, "eval_freq": 96}   # :36-46 (shortened run)
facts = {}


# ---------------------------------------------------------------- main_6DOF.py
def make_env():
    env = gym.make("my_environment/Falcon6DOF-v0", **env_config)
    env = TimeLimit(env, max_episode_steps=sb3_config["max_ep_timesteps"])
    return Monitor(env)


def make_annealed_env():
    env = gym.make("my_environment/Falcon6DOF-v0", **env_config)
    env = RewardAnnealing(env)
    env = TimeLimit(env, max_episode_steps=sb3_config["max_ep_timesteps"])
    return Monitor(env)


def make_eval_env():
    training_env = gym.make("my_environment/Falcon6DOF-v0", **env_config)
    training_env = TimeLimit(training_env, max_episode_steps=sb3_config["max_ep_timesteps"])
    return Monitor(RecordVideo(EpisodeAnalyzer(training_env), video_folder="eval_videos", episode_trigger=lambda x: x % 5 == 0))


run = wandb.init(config={**env_config, **sb3_config}, project="RL_rocket_6DOF")
env = make_env()
facts["env_class"] = type(env.unwrapped).__module__ + "." + type(env.unwrapped).__name__
model = PPO(sb3_config["policy_type"], env, tensorboard_log=f"runs/{run.id}", verbose=2, seed=env_config["seed"], ent_coef=0.01)
eval_env = DummyVecEnv([make_eval_env])
callbacks = [EvalCallback(eval_env, eval_freq=sb3_config["eval_freq"], n_eval_episodes=5, render=False, deterministic=True),
             WandbCallback(model_save_path=f"models/{run.id}", verbose=2)]
model.learn(total_timesteps=sb3_config["total_timesteps"], callback=callbacks)
phase1 = model.rollout
annealed_env = make_annealed_env()
model.set_env(annealed_env)
model.learn(total_timesteps=sb3_config["total_timesteps"], callback=callbacks)
phase2 = model.rollout
facts["phase1_steps"], facts["phase2_steps"] = int(phase1["rewards"].size), int(phase2["rewards"].size)
facts["phase1_obs_shape"] = list(phase1["obs"].shape[1:])
# the annealed reward is attitude_constraint + rew_goal - xi (a_thr + 1): bounded by 2 xi per step away from violations
xi = env_config["reward_coeff"]["xi"]
r2 = phase2["rewards"].ravel()
facts["annealed_rewards_in_range"] = bool(np.all((r2 <= 1e-12) & ((r2 >= -2 * xi - 1e-9) | (r2 <= -10 + 1e-9) | (r2 >= 10 - 2 * xi - 1e-9))))
facts["phase1_reward_differs_from_annealed"] = bool(abs(phase1["rewards"].mean() - r2.mean()) > 1e-3)
eval_obs = eval_env.reset()
facts["eval_obs_shape"] = list(eval_obs.shape)
run.finish()

# ---------------------------------------------------------------- test_6DOF_sb_integration.py
env = Rocket6DOF(**env_config)
check_env(env, skip_render_check=False)
env.close()
env = Rocket6DOF(**env_config)
model = PPO("MlpPolicy", env, verbose=1)
mean_reward, std_reward = evaluate_policy(model, env, n_eval_episodes=2, render=False)
facts["evaluate_policy_mean_finite"] = bool(np.isfinite(mean_reward))

# ---------------------------------------------------------------- test_6DOF_environment.py: a fixed action until done
kat_config = dict(env_config, ICRange=[0] * 14)             # the zero-range initial condition of test_6DOF_reward.py:10-11 (KAT-2)
env = Rocket6DOF(**kat_config)
null_action = [0.0, 0.0, -1]
done, steps, ret = False, 0, 0.0
obs = env.reset()
while not done:
    obs, rew, done, info = env.step(null_action)
    steps += 1
    ret += rew
facts["null_action_episode_steps"] = steps
facts["null_action_episode_return"] = ret
facts["null_action_final_altitude"] = float(info["state_history"][-1][0])

# ---------------------------------------------------------------- main.py (3DOF)
config = {"env_id": "my_environment/Falcon3DOF-v0", "timestep": 0.05, "max_time": 150, "RANDOM_SEED": 42}
c3 = env_config_3dof()
config.update(initial_conditions=c3["IC"], initial_conditions_range=c3["ICRange"], reward_coefficients=c3["reward_coeff"])
config["max_ep_timesteps"] = int(config["max_time"] / config["timestep"])


def make_env3():
    env = gym.make(config["env_id"], IC=config["initial_conditions"], ICRange=config["initial_conditions_range"],
                   timestep=config["timestep"], seed=config["RANDOM_SEED"], reward_coeff=config["reward_coefficients"])
    env = TimeLimit(env, max_episode_steps=config["max_ep_timesteps"])
    return Monitor(env, allow_early_resets=True, filename="logs_PPO")


env = make_env3()
model = PPO("MlpPolicy", env, seed=config["RANDOM_SEED"], ent_coef=0.01)
eval_env3 = DummyVecEnv([lambda: RecordVideoFigure(make_env3(), video_folder="videos/x", image_folder="images/x",
                                                   episode_trigger=lambda x: x % 5 == 0)])
model.learn(total_timesteps=128, callback=[EvalCallback(eval_env3, eval_freq=config["max_ep_timesteps"] / 20, n_eval_episodes=5)])
facts["obs3_dtype"] = str(model.rollout["obs"].dtype)


def make_env3_annealed(config):
    env = gym.make(config["env_id"], IC=config["initial_conditions"], ICRange=config["initial_conditions_range"],
                   timestep=config["timestep"], seed=config["RANDOM_SEED"])
    env = RewardAnnealing(env)
    return TimeLimit(env, max_episode_steps=config["max_ep_timesteps"])


model.set_env(make_env3_annealed(config))
model.learn(total_timesteps=128)
facts["phase2_3dof_steps"] = int(model.rollout["rewards"].size)

# the two commented-out compositions of main.py:57-64, :139-146 and the Gaudet observation wrapper
env = TimeLimit(DiscreteActions3DOF(gym.make(config["env_id"])), max_episode_steps=50)
m = PPO("MlpPolicy", env, seed=1)
m.learn(total_timesteps=64)
facts["discrete_action_space_n"] = int(env.action_space.n)
env = GaudetStateObs(gym.make(config["env_id"], IC=config["initial_conditions"], ICRange=config["initial_conditions_range"],
                              timestep=config["timestep"], seed=config["RANDOM_SEED"], reward_coeff=config["reward_coefficients"]))
o = env.reset()
o2, r, d, info = env.step(np.float32([0.1, 0.5]))
facts["gaudet_obs_shape"], facts["gaudet_obs_dtype"] = list(o.shape), str(o2.dtype)
facts["gaudet_altitude"] = float(o[3])
print("FACTS " + json.dumps(facts))
