"""Constructor-time constants of the environments, flattened into the POD the kernels take by value.

Mirrors what the reference constructors derive once on the host, with the same numpy expressions so
that float32/float64 promotion (numpy >= 2, "NEP 50") gives the same bits:
  Rocket6DOF.__init__  my_environment/envs/rocket_env.py:511-663
  Rocket.__init__      my_environment/envs/rocket_env.py:27-135
Default kwargs are the reference's defaults (rocket_env.py:27-43, :511-534).
"""
import numpy as np

from ._native import MAX_DISCRETE, NMAX, RlrParams

MAX_GIMBAL = np.deg2rad(20)      # rocket_env.py:66, :572
MAX_THRUST = 981e3               # rocket_env.py:67, :573

DEFAULTS_6DOF = dict(
    IC=[500, 100, 100, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 50e3],
    ICRange=[50, 10, 10, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 1e3],
    timestep=0.1,
    seed=42,
    reward_coeff={"alfa": -0.01, "beta": -1e-8, "eta": 2, "gamma": -10, "delta": -5, "kappa": 10, "xi": 0.004},
    trajectory_limits={"attitude_limit": [1.5, 1.5, 2 * np.pi]},
    landing_params={"waypoint": 50, "landing_radius": 30, "maximum_velocity": 10,
                    "landing_attitude_limit": [0.2, 0.2, 2 * np.pi], "omega_lim": [0.2, 0.2, 0.2]},
)
DEFAULTS_3DOF = dict(
    IC=[100, 500, np.pi / 2, -10, -50, 0, 50e3],
    ICRange=[10, 50, 0.1, 1, 10, 0.1, 1e3],
    timestep=0.1,
    seed=42,
    reward_coeff={"alfa": -0.01, "beta": -1e-8, "eta": 2, "gamma": -10, "delta": -5, "kappa": 10, "waypoint": 50, "landing_radius": 30},
)
DEFAULT_DISCRETE_TABLE = [[0, -1], [-1, +1], [0, +1], [+1, +1]]     # wrappers.py:25


def timestep_ms(timestep):
    ms = round(timestep * 1000.0)
    if ms <= 0 or abs(ms - timestep * 1000.0) > 1e-9:
        raise ValueError("timestep must be a positive whole number of milliseconds: the simulator rounds t to 3 decimals "
                         "(simulator.py:73, :245) and the kernels use t_k = k*dt_ms/1000")
    return int(ms)


def _angle_limit_codes(limits):
    """Limits on |euler_zyx| as the fast kernels test them without atan2/asin: for the two atan2 angles (index 0, 2)
    |angle| > L  <=>  a < cos(L) * hypot(a, b);  for the asin angle (index 1)  |angle| > L  <=>  |v| > sin(L).
    Limits outside the angle's range get sentinels that make the comparison constant."""
    out = []
    for i, L in enumerate(limits):
        L = float(L)
        if i == 1:
            out.append(-1.0 if L < 0 else (2.0 if L >= np.pi / 2 else float(np.sin(L))))
        else:
            out.append(2.0 if L < 0 else (-2.0 if L >= np.pi else float(np.cos(L))))
    return out


class EnvSpec:
    """Everything derived from the constructor kwargs: the kernel POD plus the host-side arrays."""

    def __init__(self, dof, IC=None, ICRange=None, timestep=None, seed=None, reward_coeff=None, trajectory_limits=None,
                 landing_params=None, max_episode_steps=0, reward_mode=0, discrete_table=None, obs_mode=0, min_thrust=0.0,
                 aero_mode=0):
        if dof not in (6, 3, 1):
            raise ValueError("dof must be 6, 3 or 1")
        d = DEFAULTS_6DOF if dof == 6 else DEFAULTS_3DOF
        self.dof = dof
        self.IC = d["IC"] if IC is None else IC
        self.ICRange = d["ICRange"] if ICRange is None else ICRange
        self.timestep = d["timestep"] if timestep is None else timestep
        self.seed = d["seed"] if seed is None else seed
        self.reward_coeff = d["reward_coeff"] if reward_coeff is None else reward_coeff
        self.trajectory_limits = DEFAULTS_6DOF["trajectory_limits"] if trajectory_limits is None else trajectory_limits
        self.landing_params = DEFAULTS_6DOF["landing_params"] if landing_params is None else landing_params
        self.max_episode_steps = int(max_episode_steps or 0)
        self.reward_mode = int(reward_mode)
        self.obs_mode = {0: 0, 1: 1, "state": 0, "gaudet": 1}[obs_mode]
        # extension (no reference behaviour): lower bound of the de-normalised thrust, the reference's open TODO (rocket_env.py:405, :980)
        self.min_thrust = float(min_thrust)
        if not (0.0 <= self.min_thrust <= MAX_THRUST):
            raise ValueError("min_thrust must lie in [0, max_thrust]")
        if self.obs_mode == 1 and dof != 3:
            raise ValueError("obs_mode='gaudet' (GaudetStateObs, wrappers.py:51-66) is defined for the 3DOF environment")
        # extension: the angle-of-attack line of Simulator3DOF.RHS the reference keeps commented out (simulator.py:103-104)
        self.aero_mode = {0: 0, 1: 1, False: 0, True: 1, "off": 0, "aoa": 1}[aero_mode]
        if self.aero_mode and dof == 6:
            raise ValueError("aero_mode is defined for the 3DOF / 1DOF environment (the reference's 6DOF aerodynamic force is a zero stub, "
                             "simulator.py:359-360)")
        self.ndim = 14 if dof == 6 else 7
        self.nact = {6: 3, 3: 2, 1: 1}[dof]
        self.nobs = 2 if dof == 1 else (4 if self.obs_mode == 1 else self.ndim)
        if len(self.IC) != self.ndim or len(self.ICRange) != self.ndim:
            raise ValueError(f"IC and ICRange must have {self.ndim} elements")

        self.ICMean = np.float32(self.IC)
        self.ICRangeArr = np.float32(self.ICRange)
        self.ic_low = self.ICMean - self.ICRangeArr / 2
        self.ic_high = self.ICMean + self.ICRangeArr / 2
        self.max_gimbal = MAX_GIMBAL
        self.max_thrust = MAX_THRUST
        if dof == 6:
            self._derive_6dof()
        else:
            self._derive_3dof()
        self.discrete_table = None
        self.discrete_rows = None
        self.params = RlrParams()        # ONE struct for the life of the spec: batches keep a byref to it, so updates are in place
        if discrete_table is not None:
            self.set_discrete_table(discrete_table)
        self._pack()

    # rocket_env.py:575-620
    def _derive_6dof(self):
        m = self.ICMean
        t_free_fall = (-m[3] + np.sqrt(m[3] ** 2 + 2 * 9.81 * m[0])) / 9.81
        inertia = 6.04e6
        lever_arm = 15.0
        omega_max = self.max_thrust * np.sin(self.max_gimbal) * lever_arm / (inertia) * t_free_fall / 5.0
        v_max = 2 * 9.81 * t_free_fall
        self.state_normalizer = np.maximum(
            np.array([1.2 * abs(m[0]), 1.5 * abs(m[1]), 1.5 * abs(m[2]), v_max, v_max, v_max, 1.1, 1.1, 1.1, 1.1,
                      omega_max, omega_max, omega_max, m[13] + self.ICRangeArr[13]]), 1)
        high = 0.9 * np.maximum(self.state_normalizer[0:3], 100)
        low = -0.9 * np.maximum(self.state_normalizer[1:3], 100)
        low = np.insert(low, 0, -30)
        self.pos_low = low.astype(np.float32)
        self.pos_high = high.astype(np.float32)

    # rocket_env.py:69-100
    def _derive_3dof(self):
        m = self.ICMean
        t_free_fall = (-m[4] + np.sqrt(m[4] ** 2 + 2 * 9.81 * m[1])) / 9.81
        inertia = 6.04e6
        lever_arm = 30.0
        self.state_normalizer = np.maximum(
            np.array([1.5 * abs(m[0]), 1.5 * abs(m[1]), 2 * np.pi, 2 * 9.81 * t_free_fall, 2 * 9.81 * t_free_fall,
                      self.max_thrust * np.sin(self.max_gimbal) * lever_arm / (inertia) * t_free_fall / 5.0,
                      m[6] + self.ICRangeArr[6]]), 1)
        self.x_bound_right = 0.9 * np.maximum(self.state_normalizer[0], 100)
        self.x_bound_left = -self.x_bound_right
        self.y_bound_up = 0.9 * np.maximum(self.state_normalizer[1], 100)
        self.y_bound_down = -30

    def set_discrete_table(self, table):
        """DiscreteActions3DOF rows -> physical (gimbal, thrust) f32 pairs, computed as Rocket._denormalize_action
        does for the wrapper's int64 rows (wrappers.py:34 -> rocket_env.py:395-406)."""
        if self.dof != 3:
            raise ValueError("discrete actions are defined for the 3DOF environment")
        rows = np.asarray(table)
        if rows.ndim != 2 or rows.shape[1] != 2 or not (0 < rows.shape[0] <= MAX_DISCRETE):
            raise ValueError(f"discrete table must be [n<= {MAX_DISCRETE}][2]")
        floor = self.min_thrust if self.min_thrust > 0 else -np.inf
        phys = [np.float32([r[0] * self.max_gimbal, max((r[1] + 1) / 2.0 * self.max_thrust, floor)]) for r in rows]
        self.discrete_table = np.stack(phys).astype(np.float32)
        self.discrete_rows = np.array(rows, dtype=np.float32)      # the normalised rows: RewardAnnealing sees row[1] (wrappers.py:80)
        if hasattr(self, "state_normalizer"):
            self._pack()

    def set_reward_mode(self, reward_mode):
        """Switch between the shaped reward (0) and RewardAnnealing (1) in place (main_6DOF.py:95-103): every batch built on this
        spec sees the change at its next step."""
        self.reward_mode = int(reward_mode)
        self.params.reward_mode = self.reward_mode

    def _pack(self):
        """(Re)fill self.params IN PLACE (batches hold a pointer to it)."""
        p = self.params
        p.dof, p.reward_mode, p.max_episode_steps = self.dof, self.reward_mode, self.max_episode_steps
        p.obs_mode = self.obs_mode
        p.min_thrust = self.min_thrust
        p.aero_mode = self.aero_mode
        p.dt = float(self.timestep)
        p.dt_ms = timestep_ms(self.timestep)
        rc = self.reward_coeff
        for i in range(NMAX):
            p.normalizer[i] = float(self.state_normalizer[i]) if i < self.ndim else 1.0
            p.inv_normalizer[i] = 1.0 / p.normalizer[i]
        for i in range(self.ndim):
            p.ic_low[i] = float(self.ic_low[i])
            p.ic_high[i] = float(self.ic_high[i])
        if self.dof == 6:
            lp, tl = self.landing_params, self.trajectory_limits
            p.pos_low[:] = [float(v) for v in self.pos_low]
            p.pos_high[:] = [float(v) for v in self.pos_high]
            p.waypoint, p.landing_radius, p.maximum_velocity = lp["waypoint"], lp["landing_radius"], lp["maximum_velocity"]
            p.att_traj_limit[:] = [float(v) for v in tl["attitude_limit"]]
            p.att_land_limit[:] = [float(v) for v in lp["landing_attitude_limit"]]
            p.omega_lim[:] = [0.2, 0.2, 0.2]       # rocket_env.py:656: landing_params["omega_lim"] is ignored by the reference
            p.att_traj_c[:] = _angle_limit_codes(p.att_traj_limit)
            p.att_land_c[:] = _angle_limit_codes(p.att_land_limit)
        else:
            p.x_left, p.x_right, p.y_up = float(self.x_bound_left), float(self.x_bound_right), float(self.y_bound_up)
            p.waypoint, p.landing_radius = rc["waypoint"], rc["landing_radius"]
        p.alfa, p.beta, p.eta, p.gamma, p.kappa = rc["alfa"], rc["beta"], rc["eta"], rc["gamma"], rc["kappa"]
        p.delta = rc.get("delta", 0.0)
        p.xi = rc.get("xi", 0.01)                  # RewardAnnealing default thrust_penalty (wrappers.py:69-71)
        p.max_gimbal = float(self.max_gimbal)
        p.max_thrust = float(self.max_thrust)
        if self.discrete_table is not None:
            p.n_discrete = len(self.discrete_table)
            flat = self.discrete_table.reshape(-1)
            for i, v in enumerate(flat):
                p.discrete_table[i] = float(v)
            for i, r in enumerate(self.discrete_rows):
                p.discrete_thr[i] = float(r[1])
        return p

    def with_(self, **changes):
        """A modified copy (e.g. reward_mode=1 for the annealing phase, main_6DOF.py:95-103)."""
        kw = dict(dof=self.dof, IC=self.IC, ICRange=self.ICRange, timestep=self.timestep, seed=self.seed,
                  reward_coeff=self.reward_coeff, trajectory_limits=self.trajectory_limits, landing_params=self.landing_params,
                  max_episode_steps=self.max_episode_steps, reward_mode=self.reward_mode,
                  discrete_table=None, obs_mode=self.obs_mode, min_thrust=self.min_thrust, aero_mode=self.aero_mode)
        kw.update(changes)
        spec = EnvSpec(**kw)
        if self.discrete_table is not None and "discrete_table" not in changes:
            spec.discrete_table = self.discrete_table.copy()
            spec.discrete_rows = self.discrete_rows.copy()
            spec._pack()
        return spec


# ---------------------------------------------------------------------------------------------- canonical configurations
def env_config_6dof():
    """`env_config` of the reference's configuration_file.py:4-34 (BASELINE configs 3, 4, 5)."""
    pi = np.pi
    return {
        "timestep": 0.1, "seed": 42,
        "IC": [500, 100, 100, -50, 0, 0, 1, 0, 0, 0, 0, 0, 0, 45e3],
        "ICRange": [50, 10, 10, 10, 10, 10, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 1e3],
        "reward_coeff": {"alfa": -0.01, "beta": -1e-7, "delta": -5, "eta": 0.05, "gamma": -10, "kappa": 10, "xi": 0.004},
        "trajectory_limits": {"attitude_limit": [1.5, 1.5, 2 * pi]},
        "landing_params": {"waypoint": 50, "landing_radius": 30, "maximum_velocity": 10,
                           "landing_attitude_limit": [10 / 180 * pi, 10 / 180 * pi, 2 * pi], "omega_lim": [0.2, 0.2, 0.2]},
    }


def env_config_3dof():
    """`config` of the reference's main.py:23-43 as main.py:48-56 passes it to the Rocket constructor (BASELINE config 2)."""
    pi = np.pi
    return {
        "IC": [-1600, 2000, pi * 3 / 4, 180, -90, 0, 50e3], "ICRange": [5, 50, 0, 0, 0, 0, 1e3], "timestep": 0.05, "seed": 42,
        "reward_coeff": {"alfa": -0.01, "beta": 0, "delta": -5, "eta": 0.2, "gamma": -10, "kappa": 10, "xi": 0.004,
                         "waypoint": 30, "landing_radius": 30},
    }


def env_config_1dof():
    """BASELINE config 1: the main.py configuration restricted to the vertical (README.md:4-8)."""
    return dict(IC=[0, 2000, np.pi / 2, 0, -90, 0, 50e3], ICRange=[0, 50, 0, 0, 0, 0, 1e3], timestep=0.05, seed=42,
                reward_coeff=env_config_3dof()["reward_coeff"])


MAX_EPISODE_STEPS = {6: 800, 3: 3000, 1: 3000}      # configuration_file.py:37,44 (80 s / 0.1 s); main.py:28,45 (150 s / 0.05 s)
