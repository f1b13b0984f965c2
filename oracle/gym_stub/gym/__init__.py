"""Minimal stand-in for the `gym` package (TEST INFRASTRUCTURE ONLY).

The reference (JJJerome/mbt_gym) imports `gym` for three names only: `gym.Env`,
`gym.Wrapper` and `gym.spaces.{Space,Box,MultiBinary}`.  `gym` is not
installed in this image and there is no network, so `oracle/` puts this stub on
`sys.path` to import the *unmodified* reference from /root/reference when pinning
the oracle and generating golden fixtures.  Nothing in `mbt_gym_b200/` imports it.
"""
from . import spaces  # noqa: F401


class Env:
    metadata = {}

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def seed(self, seed=None):
        return None

    def close(self):
        return None


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)
