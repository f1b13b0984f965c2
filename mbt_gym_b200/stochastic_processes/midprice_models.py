"""Midprice model descriptors (reference: mbt_gym/stochastic_processes/midprice_models.py).

Supported on the device: Constant (:12-33), BrownianMotion (:36-68), GeometricBrownianMotion (:71-111), Ou (:114-146),
BrownianMotionJump (:193-230), OuJump (:233-273), Heston (:322-372).
The reference's short-term-alpha and CEV models do not run for num_trajectories > 1 in the reference itself
 and are not built.
"""
from math import sqrt

import numpy as np

from .. import _abi
from .StochasticProcessModel import StochasticProcessModel

MidpriceModel = StochasticProcessModel


class _SymmetricBoundsMidprice(StochasticProcessModel):
    """Common ctor: observation bounds are initial_price -/+ (max - initial_price)."""

    def _finish(self, initial_price, terminal_time, step_size, num_trajectories, seed):
        hi = self._get_max_value(initial_price, terminal_time)
        super().__init__([[initial_price - (hi - initial_price)]], [[hi]], step_size, terminal_time, [[initial_price]],
                         num_trajectories, seed)

    def _flatten(self, cfg):
        cfg.midprice = self.KIND
        cfg.mid_initial = float(self.initial_state[0, 0])
        cfg.mid_step = float(self.step_size)
        cfg.mid_drift = float(getattr(self, "drift", 0.0))
        cfg.mid_vol = float(getattr(self, "volatility", 0.0))
        cfg.ou_level = float(getattr(self, "mean_reversion_level", 0.0))
        cfg.ou_speed = float(getattr(self, "mean_reversion_speed", 0.0))
        cfg.mid_jump = float(getattr(self, "jump_size", 0.0))


class ConstantMidpriceModel(_SymmetricBoundsMidprice):
    KIND = _abi.MBT_MID_CONSTANT

    def __init__(self, initial_price=100, terminal_time=1.0, step_size=0.01, num_trajectories=1, seed=None):
        self._finish(initial_price, terminal_time, step_size, num_trajectories, seed)

    def _get_max_value(self, initial_price, terminal_time):
        return initial_price


class BrownianMotionMidpriceModel(_SymmetricBoundsMidprice):
    """dS = drift dt + volatility dW; bounds S0 +/- 4 volatility sqrt(T)."""
    KIND = _abi.MBT_MID_BM

    def __init__(self, drift=0.0, volatility=2.0, initial_price=100, terminal_time=1.0, step_size=0.01,
                 num_trajectories=1, seed=None):
        self.drift, self.volatility = drift, volatility
        self._finish(initial_price, terminal_time, step_size, num_trajectories, seed)

    def _get_max_value(self, initial_price, terminal_time):
        return initial_price + 4 * self.volatility * np.sqrt(terminal_time)


class GeometricBrownianMotionMidpriceModel(_SymmetricBoundsMidprice):
    """dS = drift S dt + volatility S dW; bounds mean + 4 standard deviations of S_T."""
    KIND = _abi.MBT_MID_GBM

    def __init__(self, drift=0.0, volatility=0.1, initial_price=100, terminal_time=1.0, step_size=0.01,
                 num_trajectories=1, seed=None):
        self.drift, self.volatility = drift, volatility
        self._finish(initial_price, terminal_time, step_size, num_trajectories, seed)

    def _get_max_value(self, initial_price, terminal_time):
        var = initial_price ** 2 * np.exp(2 * self.drift * terminal_time) * (np.exp(self.volatility ** 2 * terminal_time) - 1)
        return initial_price * np.exp(self.drift * terminal_time) + 4 * sqrt(var)


class OuMidpriceModel(_SymmetricBoundsMidprice):
    """Ornstein-Uhlenbeck midprice.  NB the reference's Euler step adds -speed*(S-level) WITHOUT a dt factor
    (midprice_models.py:140-143); the kernel reproduces that literally."""
    KIND = _abi.MBT_MID_OU

    def __init__(self, mean_reversion_level=0.0, mean_reversion_speed=1.0, volatility=2.0, initial_price=100.0,
                 terminal_time=1.0, step_size=0.01, num_trajectories=1, seed=None):
        self.mean_reversion_level, self.mean_reversion_speed, self.volatility = (mean_reversion_level,
                                                                                 mean_reversion_speed, volatility)
        self._finish(initial_price, terminal_time, step_size, num_trajectories, seed)

    def _get_max_value(self, initial_price, terminal_time):
        return initial_price + 4 * self.volatility * terminal_time


class BrownianMotionJumpMidpriceModel(_SymmetricBoundsMidprice):
    """Brownian midprice that also jumps by +/- jump_size when the agent's ask / bid order is filled (:193-230).
    Only meaningful with limit-order dynamics (it needs fills)."""
    KIND = _abi.MBT_MID_BM_JUMP

    def __init__(self, drift=0.0, volatility=2.0, jump_size=1.0, initial_price=100, terminal_time=1.0, step_size=0.01,
                 num_trajectories=1, seed=None):
        self.drift, self.volatility, self.jump_size = drift, volatility, jump_size
        self._finish(initial_price, terminal_time, step_size, num_trajectories, seed)

    def _get_max_value(self, initial_price, terminal_time):
        return initial_price + 4 * self.volatility * terminal_time


class OuJumpMidpriceModel(_SymmetricBoundsMidprice):
    """OU midprice (drift not scaled by dt, as in the reference) with fill-driven jumps (:233-273)."""
    KIND = _abi.MBT_MID_OU_JUMP

    def __init__(self, mean_reversion_level=0.0, mean_reversion_speed=1.0, volatility=2.0, jump_size=1.0,
                 initial_price=100.0, terminal_time=1.0, step_size=0.01, num_trajectories=1, seed=None):
        self.mean_reversion_level, self.mean_reversion_speed = mean_reversion_level, mean_reversion_speed
        self.volatility, self.jump_size = volatility, jump_size
        self._finish(initial_price, terminal_time, step_size, num_trajectories, seed)

    def _get_max_value(self, initial_price, terminal_time):
        return initial_price + 4 * self.volatility * terminal_time


class HestonMidpriceModel(StochasticProcessModel):
    """Heston stochastic-volatility midprice (:322-372): state = (price, variance),
        S' = S + drift S dt + sqrt(v dt) S W_S,     v' = |v + rate (level - v) dt + volvol sqrt(v dt) W_v|,
    corr(W_S, W_v) = weiner_correlation.  The reference draws the pair from the GLOBAL `np.random` (its `seed` argument has no
    effect there); here it comes from the environment's Philox key like every other draw: W_S = z, W_v = rho z +
    sqrt(1 - rho^2) z2.  Like the reference, the model publishes bounds for the PRICE column only (:343-345), so the
    observation space has one entry less than the observation has columns and `normalise_observation_space=True` is
    rejected (the reference fails to broadcast)."""
    KIND = _abi.MBT_MID_HESTON

    def __init__(self, drift=0.05, volatility_mean_reversion_rate=3, volatility_mean_reversion_level=0.04,
                 weiner_correlation=-0.8, volatility_of_volatility=0.6, initial_price=100, initial_variance=0.2 ** 2,
                 terminal_time=1.0, step_size=0.01, num_trajectories=1, seed=None):
        self.drift = drift
        self.volatility_mean_reversion_rate = volatility_mean_reversion_rate
        self.weiner_correlation = weiner_correlation
        self.volatility_mean_reversion_level = volatility_mean_reversion_level
        self.volatility_of_volatility = volatility_of_volatility
        hi = self._get_max_value(initial_price, terminal_time)
        super().__init__([[initial_price - (hi - initial_price)]], [[hi]], step_size, terminal_time,
                         [[initial_price, initial_variance]], num_trajectories, seed)

    def _get_max_value(self, initial_price, terminal_time):
        return initial_price + 4 * self.volatility_mean_reversion_level * terminal_time

    def _flatten(self, cfg):
        cfg.midprice = self.KIND
        cfg.mid_initial = float(self.initial_state[0, 0])
        cfg.mid_step = float(self.step_size)
        cfg.mid_drift = float(self.drift)
        cfg.heston_speed = float(self.volatility_mean_reversion_rate)
        cfg.heston_level = float(self.volatility_mean_reversion_level)
        cfg.heston_corr = float(self.weiner_correlation)
        cfg.heston_volvol = float(self.volatility_of_volatility)
        cfg.heston_var0 = float(self.initial_state[0, 1])
