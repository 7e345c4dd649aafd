"""Single-environment gym facades: `Rocket6DOF`, `Rocket` (3DOF) and `Rocket1DOF`, same constructor kwargs, spaces,
`reset()` / `step(a)` / `seed()` surface as the reference classes (my_environment/envs/rocket_env.py:19, :505), each
backed by a batch of ONE environment on the GPU (the same kernels and C ABI the vectorised path uses).

Kept from the reference's surface: observation_space / action_space / metadata / state_names / action_names /
reward_coefficients / SIM.states|actions|times / _get_normalizer / states_to_dataframe / actions_to_dataframe /
vtarg_to_dataframe / used_mass.  Rendering (pygame / pyvista / plotly, rocket_env.py:249-383, :721-961) is out of
scope of the hot path and raises NotImplementedError.
"""
import numpy as np

from . import spaces
from .config import EnvSpec


class _History:
    """Stand-in for the `SIM` attribute wrappers reach into (simulator.py:28-31, :201-206): trajectory history lists."""

    def __init__(self, ic, nact):
        self.states = [np.array(ic)]
        self.actions = [[0] * nact]
        self.times = [0]
        self.state = np.array(ic)
        self.t = 0


class _SingleEnv(spaces.GymEnv):
    _dof = None

    def _init_common(self, spec, device, math_mode):
        from .engine import RocketBatch      # raises without the CUDA extension / a GPU (no CPU fallback)
        self._spec = spec
        self._batch = RocketBatch(spec, 1, device=device, math_mode=math_mode, debug_outputs=True)
        self.ICMean, self.ICRange = spec.ICMean, spec.ICRangeArr
        self.timestep = spec.timestep
        self.metadata = dict(self.metadata)
        self.metadata["render_fps"] = 1 / spec.timestep
        self.reward_coefficients = spec.reward_coeff
        self.init_space = spaces.Box(low=spec.ic_low, high=spec.ic_high)
        self.max_gimbal, self.max_thrust = spec.max_gimbal, spec.max_thrust
        self.state_normalizer = spec.state_normalizer
        self.observation_space = spaces.Box(low=-1, high=1, shape=(spec.nobs,))
        self.action_space = spaces.Box(low=-1, high=1, shape=(spec.nact,))
        self.SIM = None
        self.vtarg_history = []
        self.infos = []
        self._episode = 0
        self.seed(spec.seed)

    def _set_obs_mode(self, obs_mode):
        """Observation-mode switch used by wrappers.GaudetStateObs: the batch of one is rebuilt around the modified spec."""
        from .engine import RocketBatch
        self._spec = self._spec.with_(obs_mode=obs_mode)
        old = self._batch
        self._batch = RocketBatch(self._spec, 1, device=old.device, seed=old.seed, debug_outputs=True,
                                  math_mode="exact" if old.math_mode == 0 else "fast")
        self.observation_space = spaces.Box(low=-1, high=1, shape=(self._spec.nobs,))

    # ---- gym API
    def seed(self, seed=42):
        """Keys the device-side Philox IC sampler (the reference seeds gym's RandomState here, rocket_env.py:478-480)."""
        self._batch.reseed(seed)
        self.init_space.seed(seed)
        return [seed]

    def reset(self):
        self.vtarg_history = []
        obs = self._batch.reset()
        y = self._batch.y[:, 0].cpu().numpy()
        self.initial_condition = y.astype(np.float32)
        self._set_state(self.initial_condition)
        self.SIM = _History(self.initial_condition, self._spec.nact if self._dof != 1 else 2)
        return self._format_obs(obs[0].cpu().numpy(), self.initial_condition)

    def step(self, normalized_action):
        import torch
        a = self._physical_action(normalized_action)
        act = torch.as_tensor(a, dtype=torch.float32, device=self._batch.device).reshape(1, -1)
        b = self._batch
        # auto-reset is a VecEnv feature; the single env keeps its terminal state like the reference does
        k_prev, ep_prev, v0_prev, ret_prev = b.k.clone(), b.episode.clone(), b.v0.clone(), b.ep_ret.clone()
        if self._spec.reward_mode == 1:
            raise RuntimeError("use wrappers.RewardAnnealing / RocketVecEnv(reward_annealing=True) for the annealed reward")
        b.step_tensor(act, physical=True)
        y = b.y_pre_reset[:, 0].cpu().numpy()
        done = bool(b.done[0].item())
        flags = int(b.flags[0].item())
        reward = float(b.reward64[0].item())
        terms = b.reward_terms[:, 0].cpu().numpy()
        nv = 3 if self._dof == 6 else 2
        self.vtarg_history.append(b.vtarg[:nv, 0].cpu().numpy())      # rocket_env.py:245 / :1012
        if done:    # undo the fused auto-reset: restore the terminal state so that further step() calls behave like the reference
            b.y[:, 0] = b.y_pre_reset[:, 0]
            b.k.copy_(k_prev + 1)
            b.episode.copy_(ep_prev)
            b.v0.copy_(v0_prev)
            b.ep_ret.copy_(ret_prev + reward)
        self._set_state(y)
        self.action = a
        self.SIM.state = y
        self.SIM.t = round(self.SIM.t + self.timestep, 3)
        self.SIM.states.append(y)
        self.SIM.actions.append(a)
        self.SIM.times.append(self.SIM.t)
        names = ["velocity_tracking", "thrust_penalty", "eta", "attitude_constraint", "attitude_hint", "rew_goal"]
        if self._dof == 6:
            names.remove("attitude_hint")
        rewards_dict = {n: float(terms[["velocity_tracking", "thrust_penalty", "eta", "attitude_constraint", "attitude_hint",
                                        "rew_goal"].index(n)]) for n in names}
        info = {"rewards_dict": rewards_dict, "is_done": done, "state_history": self.SIM.states, "action_history": self.SIM.actions,
                "timesteps": self.SIM.times, "bounds_violation": bool(flags & 1)}
        obs = b.terminal_obs[0].cpu().numpy() if done else b.obs[0].cpu().numpy()
        return self._format_obs(obs, y.astype(np.float32)), reward, done, info

    def render(self, mode="human"):
        raise NotImplementedError("rendering is outside the hot path this package replaces (see DESIGN.md, out of scope)")

    def close(self):
        pass

    # ---- helpers shared with the reference's surface
    def _get_normalizer(self):
        return self.state_normalizer

    def _denormalize_obs(self, obs):
        return obs * self.state_normalizer

    def states_to_dataframe(self):
        import pandas as pd
        return pd.DataFrame(self.SIM.states, columns=self.state_names)

    def actions_to_dataframe(self):
        import pandas as pd
        return pd.DataFrame(self.SIM.actions, columns=self.action_names)

    def used_mass(self):
        return self.SIM.states[0][-1] - self.SIM.states[-1][-1]


class Rocket6DOF(_SingleEnv):
    """6DOF rocket landing env (reference: rocket_env.py:505-1097)."""

    metadata = {"render.modes": ["human", "rgb_array"], "render_fps": 30}
    _dof = 6

    def __init__(self, IC=None, ICRange=None, timestep=None, seed=None, reward_coeff=None, trajectory_limits=None,
                 landing_params=None, device="cuda:0", math_mode="exact", min_thrust=0.0):
        self.state_names = ["x", "y", "z", "vx", "vy", "vz", "q0", "q1", "q2", "q3", "omega1", "omega2", "omega3", "mass"]
        self.action_names = ["gimbal_y", "gimbal_z", "thrust"]
        spec = EnvSpec(6, IC, ICRange, timestep, seed, reward_coeff, trajectory_limits, landing_params, min_thrust=min_thrust)
        self._init_common(spec, device, math_mode)
        self.position_bounds_space = spaces.Box(low=spec.pos_low, high=spec.pos_high, dtype=np.float32)
        self.attitude_traj_limit = spec.trajectory_limits["attitude_limit"]
        lp = spec.landing_params
        self.target_r, self.maximum_v, self.waypoint = lp["landing_radius"], lp["maximum_velocity"], lp["waypoint"]
        self.landing_target = [0, 0, 0]
        self.landing_attitude_limit = lp["landing_attitude_limit"]
        self.omega_lim = np.array([0.2, 0.2, 0.2])
        self.state = None
        self.action = np.array([0.0, 0.0, 0.0])

    def _set_state(self, y):
        self.state = y

    def _format_obs(self, obs_f32, s_f32):
        return obs_f32                                             # rocket_env.py:963-964: float32

    def _physical_action(self, action):
        # rocket_env.py:969-981, evaluated with numpy exactly as the reference does (list -> f64, f32 ndarray -> f32 thrust)
        gimbal_y = action[0] * self.max_gimbal
        gimbal_z = action[1] * self.max_gimbal
        thrust = (action[2] + 1) / 2.0 * self.max_thrust
        if self._spec.min_thrust > 0:                             # extension: the reference's minThrust TODO (rocket_env.py:980)
            thrust = max(thrust, self._spec.min_thrust)
        return np.float32([gimbal_y, gimbal_z, thrust])

    _denormalize_action = _physical_action

    def vtarg_to_dataframe(self):
        import pandas as pd
        return pd.DataFrame(self.vtarg_history, columns=["v_x", "v_y", "v_z"])


class Rocket(_SingleEnv):
    """3DOF planar rocket landing env (reference: rocket_env.py:19-502)."""

    metadata = {"render.modes": ["human", "rgb_array"], "render_fps": 10}
    _dof = 3

    def __init__(self, IC=None, ICRange=None, timestep=None, seed=None, reward_coeff=None, device="cuda:0", math_mode="exact", min_thrust=0.0,
                 aero_mode=0):
        self.state_names = ["x", "z", "theta", "vx", "vz", "omega", "mass"]
        self.action_names = ["gimbal", "thrust"]
        spec = EnvSpec(self._dof, IC, ICRange, timestep, seed, reward_coeff, min_thrust=min_thrust, aero_mode=aero_mode)
        self._init_common(spec, device, math_mode)
        self.x_bound_right, self.x_bound_left = spec.x_bound_right, spec.x_bound_left
        self.y_bound_up, self.y_bound_down = spec.y_bound_up, spec.y_bound_down
        self.target_r, self.waypoint = spec.reward_coeff["landing_radius"], spec.reward_coeff["waypoint"]
        self.y = None
        self.action = np.array([0.0, 0.0])

    def _set_state(self, y):
        self.y = y

    def _format_obs(self, obs_f32, s_f32):
        if self._spec.obs_mode == 1:
            return obs_f32                                         # GaudetStateObs: np.float32 4-vector (wrappers.py:66)
        return s_f32 / self.state_normalizer                      # rocket_env.py:148, :175, :209-210: float64 from the f32 state

    def _physical_action(self, action):
        gimbal = action[0] * self.max_gimbal                      # rocket_env.py:395-406
        thrust = (action[1] + 1) / 2.0 * self.max_thrust
        if self._spec.min_thrust > 0:                             # extension: rocket_env.py:405
            thrust = max(thrust, self._spec.min_thrust)
        return np.float32([gimbal, thrust])

    _denormalize_action = _physical_action

    def vtarg_to_dataframe(self):
        import pandas as pd
        return pd.DataFrame(self.vtarg_history, columns=["v_x", "v_y"])


class Rocket1DOF(Rocket):
    """The 1DOF toy problem of README.md:4-8 (no code in the reference): the 3DOF env with the gimbal held at 0,
    action = normalised throttle in [-1, 1], observation = normalised (height, vertical speed)."""

    def __init__(self, IC=None, ICRange=None, timestep=None, seed=None, reward_coeff=None, device="cuda:0", math_mode="exact"):
        IC = [0, 2000, np.pi / 2, 0, -90, 0, 50e3] if IC is None else IC
        ICRange = [0, 50, 0, 0, 0, 0, 1e3] if ICRange is None else ICRange
        super().__init__(IC, ICRange, timestep, seed, reward_coeff, device, math_mode)
        self.observation_space = spaces.Box(low=-1, high=1, shape=(2,))
        self.action_space = spaces.Box(low=-1, high=1, shape=(1,))

    def _format_obs(self, obs_f32, s_f32):
        return (s_f32 / self.state_normalizer)[[1, 4]]

    def _physical_action(self, action):
        a = np.asarray(action, dtype=np.float32).reshape(-1)
        return super()._physical_action(np.float32([0.0, a[0]]))
