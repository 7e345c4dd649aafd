"""Drop-in alias of the reference's `my_environment` package (reference: my_environment/__init__.py:1-12).

`import my_environment` registers the gym ids `my_environment/Falcon3DOF-v0` and `my_environment/Falcon6DOF-v0` exactly as the
reference does, but their entry points (`my_environment.envs:Rocket`, `my_environment.envs:Rocket6DOF`) resolve to the
B200-backed classes of `rl_rocket_b200.envs`.  With this directory on `sys.path` instead of the reference's package,
`main.py`, `main_6DOF.py`, `test_6DOF_environment.py` and `test_6DOF_sb_integration.py` run unmodified
(tests/test_gpu_dropin.py executes their environment-construction code paths).
"""
from rl_rocket_b200 import _register_gym_ids as _register

_register(prefix="my_environment.envs")
