"""Price-impact descriptors for the speed-trading dynamics (reference: mbt_gym/stochastic_processes/price_impact_models.py)."""
from .. import _abi
from .StochasticProcessModel import StochasticProcessModel


class PriceImpactModel(StochasticProcessModel):
    def get_impact(self, action):
        raise NotImplementedError("price impact is evaluated inside the fused CUDA step kernel")

    @property
    def max_speed(self):
        raise NotImplementedError


class TemporaryPowerPriceImpact(PriceImpactModel):
    """execution price = S + coefficient * speed ** exponent; stateless  (:34-61)."""
    KIND = _abi.MBT_IMP_TEMP_POWER

    def __init__(self, temporary_impact_coefficient=0.01, temporary_impact_exponent=1.0, num_trajectories=1):
        self.temporary_impact_coefficient = temporary_impact_coefficient
        self.temporary_impact_exponent = temporary_impact_exponent
        super().__init__([[]], [[]], None, 0.0, [[]], num_trajectories, None)

    @property
    def max_speed(self):
        return 100.0

    def _flatten(self, cfg):
        cfg.impact = self.KIND
        cfg.imp_temp = float(self.temporary_impact_coefficient)
        cfg.imp_exponent = float(self.temporary_impact_exponent)


class TemporaryAndPermanentPriceImpact(PriceImpactModel):
    """execution price = S + k*speed + I, with the permanent part I += b*speed*dt carried as one state column
    (:64-96).  I enters the execution price only, not the midprice."""
    KIND = _abi.MBT_IMP_TEMP_PERM

    def __init__(self, temporary_impact_coefficient=0.01, permanent_impact_coefficient=0.01, n_steps=20 * 10,
                 terminal_time=1.0, num_trajectories=1):
        self.temporary_impact_coefficient = temporary_impact_coefficient
        self.permanent_impact_coefficient = permanent_impact_coefficient
        self.n_steps = n_steps
        bound = self.max_speed * terminal_time * permanent_impact_coefficient
        super().__init__([[-bound]], [[bound]], terminal_time / n_steps, 0.0, [[0]], num_trajectories, None)

    @property
    def max_speed(self):
        return 10.0

    def _flatten(self, cfg):
        cfg.impact = self.KIND
        cfg.imp_temp = float(self.temporary_impact_coefficient)
        cfg.imp_perm = float(self.permanent_impact_coefficient)
        cfg.imp_step = float(self.step_size)


class _TransientBase(PriceImpactModel):
    """Common part of the two Neuman-Voss (2022) transient-impact models: one state column Y with
    Y += (-resilience * Y + linear_kernel * speed) * dt  (:129-131,170-172)."""

    def _finish(self, n_steps, terminal_time, num_trajectories):
        self.n_steps = n_steps
        bound = self.max_speed * terminal_time * self.transient_impact_coefficient
        super().__init__([[-bound]], [[bound]], terminal_time / n_steps, 0.0, [[self.initial_transient_impact]],
                         num_trajectories, None)

    @property
    def max_speed(self):
        return 10.0

    def _flatten(self, cfg):
        cfg.impact = self.KIND
        cfg.imp_temp = float(getattr(self, "temporary_impact_coefficient", 0.0))
        cfg.imp_transient = float(self.transient_impact_coefficient)
        cfg.imp_resilience = float(self.resilience_coefficient)
        cfg.imp_kernel = float(self.linear_kernel_coefficient)
        cfg.imp_initial = float(self.initial_transient_impact)
        cfg.imp_step = float(self.step_size)


class TemporaryAndTransientPriceImpact(_TransientBase):
    """execution price = S + k*speed + kappa*Y  (:99-139)."""
    KIND = _abi.MBT_IMP_TEMP_TRANSIENT

    def __init__(self, temporary_impact_coefficient=0.01, transient_impact_coefficient=0.01, resilience_coefficient=0.01,
                 initial_transient_impact=0.01, linear_kernel_coefficient=0.01, n_steps=20 * 10, terminal_time=1.0,
                 num_trajectories=1):
        self.temporary_impact_coefficient = temporary_impact_coefficient
        self.transient_impact_coefficient = transient_impact_coefficient
        self.resilience_coefficient = resilience_coefficient
        self.initial_transient_impact = initial_transient_impact
        self.linear_kernel_coefficient = linear_kernel_coefficient
        self._finish(n_steps, terminal_time, num_trajectories)


class TransientPriceImpact(_TransientBase):
    """execution price = S + kappa*Y  (:142-179)."""
    KIND = _abi.MBT_IMP_TRANSIENT

    def __init__(self, transient_impact_coefficient=0.01, resilience_coefficient=0.01, initial_transient_impact=0.01,
                 linear_kernel_coefficient=0.01, n_steps=20 * 10, terminal_time=1.0, num_trajectories=1):
        self.transient_impact_coefficient = transient_impact_coefficient
        self.resilience_coefficient = resilience_coefficient
        self.initial_transient_impact = initial_transient_impact
        self.linear_kernel_coefficient = linear_kernel_coefficient
        self._finish(n_steps, terminal_time, num_trajectories)
