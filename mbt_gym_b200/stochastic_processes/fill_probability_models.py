"""Fill-probability descriptors (reference: mbt_gym/stochastic_processes/fill_probability_models.py)."""
import numpy as np

from .. import _abi
from .StochasticProcessModel import StochasticProcessModel


class FillProbabilityModel(StochasticProcessModel):
    def get_fills(self, depths):
        raise NotImplementedError("fills are drawn inside the fused CUDA step kernel (Philox), not on the host")

    @property
    def max_depth(self):
        raise NotImplementedError


class ExponentialFillFunction(FillProbabilityModel):
    """P(fill | depth) = exp(-fill_exponent * depth)  (:57-58); a depth < 0 gives a 'probability' > 1 = always filled.
    `max_depth` is the depth at which the fill probability is 1 %  (:60-62)."""
    KIND = _abi.MBT_FILL_EXPONENTIAL

    def __init__(self, fill_exponent=1.5, step_size=0.1, num_trajectories=1, seed=None):
        self.fill_exponent = fill_exponent
        super().__init__([[]], [[]], step_size, 0.0, [[]], num_trajectories, seed)

    def _get_fill_probabilities(self, depths):
        """Host helper for analysis / plotting only (the kernel evaluates its own bit-reproducible exp)."""
        return np.exp(-self.fill_exponent * np.asarray(depths))

    @property
    def max_depth(self):
        return -np.log(0.01) / self.fill_exponent

    def _flatten(self, cfg):
        cfg.fill = self.KIND
        cfg.fill_exponent = float(self.fill_exponent)
