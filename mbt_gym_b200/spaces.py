"""Observation / action spaces.

`gym` / `gymnasium` are optional: when one is importable its Box / MultiBinary are used (so stable-baselines3
accepts the spaces unchanged); otherwise the small stand-ins below provide the attributes the environment, the
agents and SB3-style callers read (`low`, `high`, `shape`, `dtype`, `sample()`, `seed()`, `contains()`).
"""
import numpy as np

try:  # pragma: no cover - depends on the installation
    from gymnasium.spaces import Box, MultiBinary, Space  # type: ignore

    BACKEND = "gymnasium"
except Exception:  # noqa: BLE001
    try:  # pragma: no cover
        from gym.spaces import Box, MultiBinary, Space  # type: ignore

        BACKEND = "gym"
    except Exception:  # noqa: BLE001
        BACKEND = "builtin"

        class Space:
            def __init__(self, shape=None, dtype=None):
                self.shape = None if shape is None else tuple(shape)
                self.dtype = None if dtype is None else np.dtype(dtype)
                self._np_random = np.random.default_rng()

            def seed(self, seed=None):
                self._np_random = np.random.default_rng(seed)
                return [seed]

            def __repr__(self):
                return f"{type(self).__name__}{self.shape}"

        class Box(Space):
            def __init__(self, low, high, shape=None, dtype=np.float32):
                if shape is None:
                    shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
                shape = tuple(shape)
                self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape).copy()
                self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape).copy()
                super().__init__(shape, dtype)

            def sample(self):
                return self._np_random.uniform(self.low, self.high, size=self.shape).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

            def __repr__(self):
                return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"

        class MultiBinary(Space):
            def __init__(self, n):
                self.n = n
                super().__init__((n,), np.int8)

            def sample(self):
                return self._np_random.integers(0, 2, size=self.shape).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all((x == 0) | (x == 1)))
