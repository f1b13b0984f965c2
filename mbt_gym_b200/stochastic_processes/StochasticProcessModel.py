"""Base class of the market-model DESCRIPTORS.

In the reference a `StochasticProcessModel` owns a numpy Generator and advances its own (N, d) state array
(mbt_gym/stochastic_processes/StochasticProcessModel.py:8-53).  Here the arithmetic lives in the CUDA step kernel;
the Python objects only carry the constructor parameters (same keyword names and attribute names, so agents such as
AvellanedaStoikovAgent can keep reading `volatility`, `intensity`, `fill_exponent` ...), the bounds that define the
observation space, and `_flatten(cfg)`, which writes the parameters into the `mbt_config` POD of the C ABI.

User subclasses with their own `update()` cannot be fused into the kernel; the environment rejects unknown process
types with NotImplementedError instead of silently running them on the CPU.
"""
import numpy as np

from .._track import Tracked


def _row(values):
    """(1, d) float array; d may be 0 for stateless processes."""
    return np.asarray(values, dtype=float).reshape(1, -1)


class StochasticProcessModel(Tracked):
    #: MBT_* enum value written to mbt_config by the environment (set by concrete classes)
    KIND = None

    def __init__(self, min_value, max_value, step_size, terminal_time, initial_state, num_trajectories=1, seed=None):
        self.min_value, self.max_value, self.initial_state = _row(min_value), _row(max_value), _row(initial_state)
        for name in ("initial_state", "min_value", "max_value"):
            a = getattr(self, name)
            assert a.ndim == 2 and a.shape[0] == 1, f"Attribute {name} must be a vector of shape (1, state_size)."
        self.step_size = step_size
        self.terminal_time = terminal_time
        self.num_trajectories = num_trajectories
        self.seed_ = seed
        self._env = None     # set by TradingEnvironment: lets `current_state` read the device state
        self._columns = None

    # -- reference surface
    @property
    def initial_vector_state(self):
        return np.repeat(self.initial_state, self.num_trajectories, axis=0)

    @property
    def current_state(self):
        """(N, d) slice of the environment state for this process (device -> host copy on access)."""
        if self._env is None or self._columns is None or not self._env._started:
            return self.initial_vector_state
        lo, hi = self._columns
        return self._env.state[:, lo:hi]

    def reset(self):
        """State is reset by TradingEnvironment.reset() on the device; nothing to do on the host."""

    def seed(self, seed=None):
        """Kept for API compatibility: randomness is keyed by the environment's Philox seed."""
        self.seed_ = seed

    def update(self, arrivals, fills, action, state=None):
        raise NotImplementedError(
            "process updates run inside the fused CUDA step kernel (libmbt_b200); there is no host-side update()")

    # -- flattening
    def _flatten(self, cfg):
        raise NotImplementedError(f"{type(self).__name__} has no CUDA implementation")
