"""Order-type descriptors: which processes a market needs, the action space, and the dynamics enum of the C ABI
(reference: mbt_gym/gym/ModelDynamics.py).  `update_state` / `get_arrivals_and_fills` of the reference are the
bodies of the fused CUDA step kernel (mbt_gym_b200/csrc/mbt_step_core.cuh)."""
import numpy as np

from .. import _abi
from .._track import Tracked
from ..spaces import Box, MultiBinary


class ModelDynamics(Tracked):
    KIND = None
    REQUIRED = ()
    round_initial_inventory = False

    def __init__(self, midprice_model=None, arrival_model=None, fill_probability_model=None, price_impact_model=None,
                 num_trajectories=1, seed=None):
        self.midprice_model = midprice_model
        self.arrival_model = arrival_model
        self.fill_probability_model = fill_probability_model
        self.price_impact_model = price_impact_model
        self.num_trajectories = num_trajectories
        self.seed_ = seed
        self.required_processes = self.get_required_stochastic_processes()
        for name in self.required_processes:
            assert getattr(self, name) is not None, f"This model dynamics cannot have env.{name} to be None."
        self._env = None

    # -- state lives on the device; these two keep `env.model_dynamics.state` readable / assignable
    @property
    def state(self):
        return None if self._env is None else self._env._get_state()

    @state.setter
    def state(self, value):
        if self._env is not None and value is not None:
            self._env._set_state(value)

    @property
    def midprice(self):
        return self.midprice_model.current_state[:, 0].reshape(-1, 1)

    @property
    def fill_multiplier(self):
        ones = np.ones((self.num_trajectories, 1))
        return np.append(-ones, ones, axis=1)

    def get_required_stochastic_processes(self):
        return list(self.REQUIRED)

    def get_action_space(self):
        raise NotImplementedError

    def _get_max_depth(self):
        return None if self.fill_probability_model is None else self.fill_probability_model.max_depth

    def _get_max_speed(self):
        return None if self.price_impact_model is None else self.price_impact_model.max_speed

    def _flatten(self, cfg):
        cfg.dynamics = self.KIND
        cfg.half_spread = float(getattr(self, "fixed_market_half_spread", 0.0))


class LimitOrderModelDynamics(ModelDynamics):
    """The agent posts a bid and an ask at depths (action[:,0], action[:,1]) from the midprice   (:87-131)."""
    KIND = _abi.MBT_DYN_LIMIT
    REQUIRED = ("arrival_model", "fill_probability_model")
    round_initial_inventory = True

    def __init__(self, midprice_model=None, arrival_model=None, fill_probability_model=None, num_trajectories=1,
                 seed=None, max_depth=None):
        super().__init__(midprice_model, arrival_model, fill_probability_model, None, num_trajectories, seed)
        self.max_depth = max_depth or self._get_max_depth()

    def get_action_space(self):
        assert self.max_depth is not None, "For limit orders max_depth cannot be None."
        return Box(low=np.float32(0.0), high=np.float32(self.max_depth), shape=(2,))


class AtTheTouchModelDynamics(ModelDynamics):
    """The agent decides, per side, whether to post at the touch (fixed half spread)   (:134-176)."""
    KIND = _abi.MBT_DYN_AT_TOUCH
    REQUIRED = ("arrival_model",)
    round_initial_inventory = True

    def __init__(self, midprice_model=None, arrival_model=None, fill_probability_model=None, num_trajectories=1,
                 fixed_market_half_spread=0.5, seed=None):
        super().__init__(midprice_model, arrival_model, fill_probability_model, None, num_trajectories, seed)
        self.fixed_market_half_spread = fixed_market_half_spread

    def get_action_space(self):
        return MultiBinary(2)


class LimitAndMarketOrderModelDynamics(ModelDynamics):
    """Limit-order depths plus market-order buy / sell switches (action[:,2:4] > 0.5)   (:179-240)."""
    KIND = _abi.MBT_DYN_LIMIT_AND_MARKET
    REQUIRED = ("arrival_model", "fill_probability_model")
    round_initial_inventory = True

    def __init__(self, midprice_model=None, arrival_model=None, fill_probability_model=None, num_trajectories=1,
                 seed=None, max_depth=None, fixed_market_half_spread=0.5):
        super().__init__(midprice_model, arrival_model, fill_probability_model, None, num_trajectories, seed)
        self.max_depth = max_depth or self._get_max_depth()
        self.fixed_market_half_spread = fixed_market_half_spread

    def get_action_space(self):
        assert self.max_depth is not None, "For limit orders max_depth cannot be None."
        return Box(low=np.zeros(4, dtype=np.float32), high=np.array([self.max_depth, self.max_depth, 1, 1], dtype=np.float32))


class TradinghWithSpeedModelDynamics(ModelDynamics):
    """The agent chooses a trading speed (positive buys); execution price = midprice + impact   (:243-275).
    (Class name spelled as in the reference.)"""
    KIND = _abi.MBT_DYN_SPEED
    REQUIRED = ("price_impact_model",)
    round_initial_inventory = False

    def __init__(self, midprice_model=None, price_impact_model=None, num_trajectories=1, seed=None, max_speed=None):
        super().__init__(midprice_model, None, None, price_impact_model, num_trajectories, seed)
        self.max_speed = max_speed or self._get_max_speed()

    def get_action_space(self):
        return Box(low=np.float32([-self.max_speed]), high=np.float32([self.max_speed]))
