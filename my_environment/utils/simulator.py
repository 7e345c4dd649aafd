"""Name-level stand-ins for `Simulator3DOF` / `Simulator6DOF` (reference: utils/simulator.py:12, :179).  The reference's
environments own one simulator object each; here the integration (scipy RK45 + ground event, restated in CUDA) is fused into
the environment step, so constructing one of these directly is an error with a pointer to the environment classes."""


class _Fused:
    def __init__(self, *a, **k):
        raise NotImplementedError("the simulator is fused into the CUDA environment step: use my_environment.envs.Rocket / Rocket6DOF "
                                  "(their SIM attribute keeps the states / actions / times history)")


class Simulator3DOF(_Fused):
    pass


class Simulator6DOF(_Fused):
    pass
