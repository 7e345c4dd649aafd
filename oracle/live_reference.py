"""Loader for the LIVE reference (Tuxliri/RL_rocket) -- TEST INFRASTRUCTURE ONLY.

Runs the reference's own, unmodified Python sources from /root/reference with
import stubs for the four packages that are absent from this image (gym,
matplotlib, pyvista, pygame; see oracle/stubs/).  This only works in the
development container: /root/reference does not exist on the GPU box, so this
module is used by `oracle/make_golden.py` (which writes tests/golden/*) and by
CPU-side tests that skip when the reference is absent.  Nothing in the product
package (`rl_rocket_b200/`) may import it.

Reference entry points reached through here:
  my_environment/utils/simulator.py:12   Simulator3DOF
  my_environment/utils/simulator.py:179  Simulator6DOF
  my_environment/envs/rocket_env.py:19   Rocket (3DOF)
  my_environment/envs/rocket_env.py:505  Rocket6DOF
  my_environment/wrappers/wrappers.py:24,68  DiscreteActions3DOF, RewardAnnealing
"""
import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get("RL_ROCKET_REFERENCE", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")
_loaded = {}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "my_environment", "utils", "simulator.py"))


def load():
    """Import the reference package (once) and return a namespace of its classes."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError(f"live reference not present under {REFERENCE_ROOT}")
    for p in (_STUBS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    os.environ.setdefault("WANDB_MODE", "disabled")
    import importlib

    sim = importlib.import_module("my_environment.utils.simulator")
    envs = importlib.import_module("my_environment.envs.rocket_env")
    try:
        wr = importlib.import_module("my_environment.wrappers.wrappers")
    except Exception:  # wandb / pandas plotting backend problems must not block the step path
        wr = None
    import gym  # the stub

    class NS:
        pass

    ns = NS()
    ns.Simulator3DOF = sim.Simulator3DOF
    ns.Simulator6DOF = sim.Simulator6DOF
    ns.Rocket = envs.Rocket
    ns.Rocket6DOF = envs.Rocket6DOF
    ns.wrappers = wr
    ns.gym = gym
    _loaded["ns"] = ns
    return ns


# --- canonical configurations (values restated from the reference's config files) -------------

def env_config_6dof():
    """configuration_file.py:4-34 (`env_config`)."""
    from math import pi

    return {
        "timestep": 0.1,
        "seed": 42,
        "IC": [500, 100, 100, -50, 0, 0, 1, 0, 0, 0, 0, 0, 0, 45e3],
        "ICRange": [50, 10, 10, 10, 10, 10, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 0.1, 1e3],
        "reward_coeff": {"alfa": -0.01, "beta": -1e-7, "delta": -5, "eta": 0.05, "gamma": -10, "kappa": 10, "xi": 0.004},
        "trajectory_limits": {"attitude_limit": [1.5, 1.5, 2 * pi]},
        "landing_params": {
            "waypoint": 50,
            "landing_radius": 30,
            "maximum_velocity": 10,
            "landing_attitude_limit": [10 / 180 * pi, 10 / 180 * pi, 2 * pi],
            "omega_lim": [0.2, 0.2, 0.2],
        },
    }


def env_config_3dof():
    """main.py:23-43 (`config`) mapped to the Rocket ctor kwargs as main.py:48-56 does."""
    from math import pi

    return {
        "IC": [-1600, 2000, pi * 3 / 4, 180, -90, 0, 50e3],
        "ICRange": [5, 50, 0, 0, 0, 0, 1e3],
        "timestep": 0.05,
        "seed": 42,
        "reward_coeff": {"alfa": -0.01, "beta": 0, "delta": -5, "eta": 0.2, "gamma": -10, "kappa": 10, "xi": 0.004,
                         "waypoint": 30, "landing_radius": 30},
    }
