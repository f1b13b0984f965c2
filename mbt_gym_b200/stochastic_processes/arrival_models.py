"""Arrival model descriptors (reference: mbt_gym/stochastic_processes/arrival_models.py).

Column 0 of every two-sided quantity is the exogenous SELL order hitting the agent's bid, column 1 the BUY order
lifting the ask (:10-13)."""
import numpy as np

from .. import _abi
from .StochasticProcessModel import StochasticProcessModel


class ArrivalModel(StochasticProcessModel):
    def get_arrivals(self):
        raise NotImplementedError("arrivals are drawn inside the fused CUDA step kernel (Philox), not on the host")


class _StatelessPoisson(ArrivalModel):
    def __init__(self, intensity=np.array([140.0, 140.0]), step_size=0.001, num_trajectories=1, seed=None):
        self.intensity = np.array(intensity)
        super().__init__([[]], [[]], step_size, 0.0, [[]], num_trajectories, seed)

    def _flatten(self, cfg):
        cfg.arrival = self.KIND
        cfg.arr_step = float(self.step_size)
        cfg.arr_rate[0], cfg.arr_rate[1] = (float(x) for x in np.asarray(self.intensity, float).reshape(-1)[:2])


class PoissonArrivalModel(_StatelessPoisson):
    """P(arrival in a step) = intensity * step_size   (:54-56)."""
    KIND = _abi.MBT_ARR_POISSON


class PoissonArrivalNonLinearModel(_StatelessPoisson):
    """P(arrival in a step) = 1 - exp(-intensity * step_size)   (:81-83)."""
    KIND = _abi.MBT_ARR_POISSON_NONLINEAR


class HawkesArrivalModel(ArrivalModel):
    """Self-exciting intensities: lambda += speed*(baseline-lambda)*dt + jump*arrival   (:110-123).
    State = (lambda_bid, lambda_ask), bounded by [0, 10*baseline] in the observation space (:125-126)."""
    KIND = _abi.MBT_ARR_HAWKES

    def __init__(self, baseline_arrival_rate=np.array([[10.0, 10.0]]), step_size=0.01, jump_size=40.0,
                 mean_reversion_speed=60.0, terminal_time=1, num_trajectories=1, seed=None):
        self.baseline_arrival_rate = baseline_arrival_rate
        self.jump_size, self.mean_reversion_speed = jump_size, mean_reversion_speed
        super().__init__([[0, 0]], np.array([[1, 1]]) * self._get_max_arrival_rate(), step_size, terminal_time,
                         baseline_arrival_rate, num_trajectories, seed)

    def _get_max_arrival_rate(self):
        return np.asarray(self.baseline_arrival_rate) * 10

    def _flatten(self, cfg):
        cfg.arrival = self.KIND
        cfg.arr_step = float(self.step_size)
        cfg.arr_rate[0], cfg.arr_rate[1] = (float(x) for x in np.asarray(self.baseline_arrival_rate, float).reshape(-1)[:2])
        cfg.hawkes_jump, cfg.hawkes_speed = float(self.jump_size), float(self.mean_reversion_speed)
