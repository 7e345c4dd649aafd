"""gym.spaces shim.  When gym (0.21) is importable its own Box/Discrete are used, so SB3's isinstance checks pass;
otherwise a minimal stand-in with the attributes SB3 / the reference scripts read (low, high, shape, dtype, n,
sample, contains, seed)."""
import numpy as np

try:  # pragma: no cover - gym is absent from the build image
    from gym import Env as GymEnv
    from gym.spaces import Box, Discrete
    HAVE_GYM = True
except Exception:  # noqa: BLE001
    HAVE_GYM = False

    class GymEnv:
        metadata = {"render.modes": []}
        reward_range = (-float("inf"), float("inf"))
        spec = None

        @property
        def unwrapped(self):
            return self

        def seed(self, seed=None):
            return [seed]

        def close(self):
            pass

    class _Space:
        def __init__(self, shape, dtype):
            self.shape = tuple(shape)
            self.dtype = np.dtype(dtype)
            self.np_random = np.random.RandomState()

        def seed(self, seed=None):
            self.np_random = np.random.RandomState(None if seed is None else int(seed) % (2 ** 32))
            return [seed]

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.shape(low) if not np.isscalar(low) else np.shape(high)
            super().__init__(shape, dtype)
            self.low = np.full(self.shape, low, dtype=self.dtype) if np.isscalar(low) else np.asarray(low).astype(self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype) if np.isscalar(high) else np.asarray(high).astype(self.dtype)

        def sample(self):
            return self.np_random.uniform(self.low, self.high, self.shape).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return bool(np.can_cast(x.dtype, self.dtype) and x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Discrete(_Space):
        def __init__(self, n):
            super().__init__((), np.int64)
            self.n = int(n)

        def sample(self):
            return int(self.np_random.randint(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

        def __repr__(self):
            return f"Discrete({self.n})"
