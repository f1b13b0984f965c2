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


class TriangularFillFunction(FillProbabilityModel):
    """Reference :68-91, implemented as written there: `np.max(1 - np.max(depths, 0) / max_fill_depth, 0)` reduces over
    the TRAJECTORY axis (and then over the two sides), so every quote of a step is filled with ONE probability that
    follows the deepest quote of the batch.  On the device this is a batch reduction in front of the step kernel
    (`mbt_fill_batch_kernel`).  `max_depth = 1.5 * max_fill_depth` (:84-86)."""
    KIND = _abi.MBT_FILL_TRIANGULAR

    def __init__(self, max_fill_depth=1.0, step_size=0.1, num_trajectories=1, seed=None):
        self.max_fill_depth = max_fill_depth
        super().__init__([[]], [[]], step_size, 0.0, [[]], num_trajectories, seed)

    def _get_fill_probabilities(self, depths):
        """Host helper for analysis only; same expression as the reference (a scalar)."""
        return np.max(1 - np.max(np.asarray(depths), 0) / self.max_fill_depth, 0)

    @property
    def max_depth(self):
        return 1.5 * self.max_fill_depth

    def _flatten(self, cfg):
        cfg.fill = self.KIND
        cfg.fill_max_depth = float(self.max_fill_depth)


class PowerFillFunction(FillProbabilityModel):
    """Reference :94-123, implemented as written there: `(1 + (fill_multiplier * np.max(depths, 0)) ** fill_exponent) ** -1`
    with `np.max(depths, 0)` over the TRAJECTORY axis -- one fill probability per side for the whole batch.
    `max_depth = 0.01 ** (-1 / fill_exponent) - 1` (:115-117)."""
    KIND = _abi.MBT_FILL_POWER

    def __init__(self, fill_exponent=1.5, fill_multiplier=1.5, step_size=0.1, num_trajectories=1, seed=None):
        self.fill_exponent = fill_exponent
        self.fill_multiplier = fill_multiplier
        super().__init__([[]], [[]], step_size, 0.0, [[]], num_trajectories, seed)

    def _get_fill_probabilities(self, depths):
        """Host helper for analysis only; same expression as the reference (one value per side)."""
        return (1 + (self.fill_multiplier * np.max(np.asarray(depths), 0)) ** self.fill_exponent) ** -1

    @property
    def max_depth(self):
        return 0.01 ** (-1 / self.fill_exponent) - 1

    def _flatten(self, cfg):
        cfg.fill = self.KIND
        cfg.fill_exponent = float(self.fill_exponent)
        cfg.fill_multiplier = float(self.fill_multiplier)


class ExogenousMmFillProbabilityModel(FillProbabilityModel):
    """Reference :126-170: beyond the exogenous best depth d of the side the fill probability decays as
    `base_fill_probability * exp(-fill_exponent * (depth - d))`, at or inside it the order is always filled.  The model owns
    two state columns, (d_bid, d_ask), initialised from the two `exogenous_best_depth_processes`.  In the reference those
    columns never change: `update()` advances the two processes but never copies their state back (:168-170), so the depths
    stay at the processes' initial values -- which is what the kernel implements (two constant observation columns).
    The processes are used as descriptors only: initial state and bounds."""
    KIND = _abi.MBT_FILL_EXOGENOUS_MM

    def __init__(self, exogenous_best_depth_processes, fill_exponent=1.5, base_fill_probability=1.0, step_size=0.1,
                 num_trajectories=1, seed=None):
        assert len(exogenous_best_depth_processes) == 2, "exogenous_best_depth_processes must be length 2 (bid and ask)"
        assert all(p.initial_state.shape[1] > 0 for p in exogenous_best_depth_processes), \
            "Exogenous best depth processes must have a state of at least size 1."
        if any(p.initial_state.shape[1] != 1 for p in exogenous_best_depth_processes):
            raise NotImplementedError("exogenous best-depth processes with more than one state column have no CUDA "
                                      "implementation (the reference cannot compare their state with the depths either)")
        self.exogenous_best_depth_processes = tuple(exogenous_best_depth_processes)
        self.fill_exponent = fill_exponent
        self.base_fill_probability = base_fill_probability
        bid, ask = self.exogenous_best_depth_processes
        super().__init__(np.concatenate([bid.min_value, ask.min_value], axis=1),
                         np.concatenate([bid.max_value, ask.max_value], axis=1), step_size, 0.0,
                         np.concatenate([bid.initial_state, ask.initial_state], axis=1), num_trajectories, seed)

    def _get_fill_probabilities(self, depths):
        """Host helper for analysis only; same expression as the reference with the (constant) exogenous depths."""
        depths, best = np.asarray(depths), self.initial_state
        return (depths > best) * self.base_fill_probability * np.exp(-self.fill_exponent * (depths - best)) + (depths <= best)

    @property
    def max_depth(self):
        return -np.log(0.01) / self.fill_exponent + np.max(self.exogenous_best_depth_processes[0].max_value)

    def _flatten(self, cfg):
        cfg.fill = self.KIND
        cfg.fill_exponent = float(self.fill_exponent)
        cfg.fill_base = float(self.base_fill_probability)
        cfg.fill_depth0[0], cfg.fill_depth0[1] = (float(x) for x in self.initial_state[0])
