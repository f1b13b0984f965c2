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
