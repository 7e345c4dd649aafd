"""`my_environment.utils` (reference: utils/simulator.py).  The simulators are not separate host objects here: the integration
runs inside the step kernels (rl_rocket_b200/csrc).  `simulator` exposes the names for code that only imports them."""
from my_environment.utils import simulator  # noqa: F401
