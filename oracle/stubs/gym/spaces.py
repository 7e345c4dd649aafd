import hashlib
import numpy as np


def _hash_seed(seed, max_bytes=8):
    h = hashlib.sha512(str(seed).encode("utf8")).digest()
    return int.from_bytes(h[:max_bytes], "little")


def _int_list_from_bigint(b):
    if b == 0:
        return [0]
    out = []
    while b > 0:
        b, m = divmod(b, 2 ** 32)
        out.append(m)
    return out


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self.np_random = np.random.RandomState()

    def seed(self, seed=None):
        if seed is None:
            seed = 0
        self.np_random = np.random.RandomState()
        self.np_random.seed(_int_list_from_bigint(_hash_seed(seed)))
        return [seed]


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        dtype = np.dtype(dtype)
        if shape is None:
            shape = np.shape(low) if not np.isscalar(low) else np.shape(high)
        shape = tuple(shape)
        self.low = np.full(shape, low, dtype=dtype) if np.isscalar(low) else np.asarray(low).astype(dtype)
        self.high = np.full(shape, high, dtype=dtype) if np.isscalar(high) else np.asarray(high).astype(dtype)
        super().__init__(shape, dtype)

    def sample(self):
        return self.np_random.uniform(low=self.low, high=self.high, size=self.shape).astype(self.dtype)

    def contains(self, x):
        if not isinstance(x, np.ndarray):
            x = np.asarray(x)
        return bool(
            np.can_cast(x.dtype, self.dtype)
            and x.shape == self.shape
            and np.all(x >= self.low)
            and np.all(x <= self.high)
        )


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        super().__init__((), np.int64)

    def sample(self):
        return int(self.np_random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n
