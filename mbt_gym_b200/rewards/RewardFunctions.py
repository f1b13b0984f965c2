"""Reward-function descriptors (reference: mbt_gym/rewards/RewardFunctions.py).

Inside `TradingEnvironment.step()` the reward is computed by the fused CUDA step kernel.  `calculate()` is kept for
callers that evaluate a reward on explicit state matrices (the reference's unit tests do): it runs the SAME device
code through `mbt_reward_eval` of the C ABI on a small private handle -- there is no NumPy implementation here.
"""
import numpy as np

from .. import _abi
from .._track import Tracked
from ..gym.index_names import INVENTORY_INDEX, TIME_INDEX


class RewardFunction(Tracked):
    KIND = None

    def __init__(self):
        self.initial_inventory = None
        self.episode_length = None
        self._handles = {}

    # -- flattening
    def _flatten(self, cfg):
        cfg.reward = self.KIND
        cfg.rew_phi = float(getattr(self, "per_step_inventory_aversion", 0.0))
        cfg.rew_alpha = float(getattr(self, "terminal_inventory_aversion", 0.0))
        cfg.rew_exponent = float(getattr(self, "inventory_exponent", 2.0))
        cfg.rew_risk_aversion = float(getattr(self, "risk_aversion", 0.0))
        if getattr(self, "terminal_time", None) is not None:
            cfg.rew_terminal_time = float(self.terminal_time)

    # -- reference surface
    def reset(self, initial_state):
        """CjMm / CjOe capture the initial inventory and the episode length (:72-74,111-113)."""
        initial_state = np.asarray(initial_state)
        self.initial_inventory = initial_state[:, INVENTORY_INDEX]
        if getattr(self, "terminal_time", None) is not None:
            self.episode_length = self.terminal_time - initial_state[:, TIME_INDEX]

    def calculate(self, current_state, action, next_state, is_terminal_step=False):
        from .. import _lib

        cur = np.asarray(current_state, dtype=np.float64)
        nxt = np.asarray(next_state, dtype=np.float64)
        assert cur.ndim > 1, "Reward functions must be calculated on state matrices."
        act = np.asarray(action, dtype=np.float64).reshape(cur.shape[0], -1)
        n, d = cur.shape
        q0 = 0.0
        ep_len = 1.0
        if self.initial_inventory is not None:
            q0s = np.unique(np.asarray(self.initial_inventory, float))
            if q0s.size != 1:
                raise NotImplementedError("calculate() on explicit matrices supports one initial inventory per call")
            q0 = float(q0s[0])
        if self.episode_length is not None:
            ep_len = float(np.asarray(self.episode_length, float).reshape(-1)[0])
        key = (n, d, act.shape[1], q0, ep_len)
        env = self._handles.get(key)
        if env is None:
            cfg = _abi.new_config(precision=_abi.MBT_F64, num_trajectories=n, n_steps=1, terminal_time=1.0,
                                  step_size=1.0, max_inventory=1e300, max_cash=1e300, q0_const=q0, mid_step=1.0,
                                  arr_step=1.0, imp_step=1.0, fill_exponent=1.0)
            if act.shape[1] == 1:
                cfg.dynamics, cfg.impact = _abi.MBT_DYN_SPEED, (_abi.MBT_IMP_TEMP_PERM if d == 5 else _abi.MBT_IMP_TEMP_POWER)
                cfg.midprice = _abi.MBT_MID_BM
            else:
                cfg.dynamics = _abi.MBT_DYN_LIMIT if act.shape[1] == 2 else _abi.MBT_DYN_LIMIT_AND_MARKET
                cfg.midprice, cfg.fill = _abi.MBT_MID_BM, _abi.MBT_FILL_EXPONENTIAL
                cfg.arrival = _abi.MBT_ARR_HAWKES if d == 6 else _abi.MBT_ARR_POISSON
            self._flatten(cfg)
            # episode length L = rew_terminal_time - start_time: encode it with start_time = 0
            cfg.rew_terminal_time = ep_len
            env = _lib.NativeEnv(cfg)
            if env.D != d:
                raise ValueError(f"state matrices have {d} columns; no supported model layout matches")
            env.reset()
            self._handles = {key: env}
        return env.reward_eval(cur, act, nxt, bool(np.asarray(is_terminal_step).reshape(-1)[0]))


class PnL(RewardFunction):
    """Mark-to-market profit of the step: (cash' + q' S') - (cash + q S)   (:20-36)."""
    KIND = _abi.MBT_REW_PNL

    def reset(self, initial_state):
        pass


class _InventoryAversion(RewardFunction):
    def __init__(self, per_step_inventory_aversion=0.01, terminal_inventory_aversion=0.0, inventory_exponent=2.0,
                 terminal_time=None):
        super().__init__()
        self.per_step_inventory_aversion = per_step_inventory_aversion
        self.terminal_inventory_aversion = terminal_inventory_aversion
        self.inventory_exponent = inventory_exponent
        self.terminal_time = terminal_time
        self.pnl = PnL()


class CjOeCriterion(_InventoryAversion):
    """Optimal-execution criterion: PnL - dt*phi*q'^p - dt*alpha*(p*nu*q^(p-1) + q0^p*L)   (:39-74)."""
    KIND = _abi.MBT_REW_CJ_OE

    def __init__(self, per_step_inventory_aversion=0.01, terminal_inventory_aversion=0.0, inventory_exponent=2.0,
                 terminal_time=1.0):
        super().__init__(per_step_inventory_aversion, terminal_inventory_aversion, inventory_exponent, terminal_time)


class CjMmCriterion(_InventoryAversion):
    """Market-making criterion with the terminal penalty spread over the path by Ito's lemma:
    PnL - dt*phi*q'^p - alpha*(q'^p - q^p + dt/L*q0^p)   (:77-113)."""
    KIND = _abi.MBT_REW_CJ_MM

    def __init__(self, per_step_inventory_aversion=0.01, terminal_inventory_aversion=0.0, inventory_exponent=2.0,
                 terminal_time=1.0):
        super().__init__(per_step_inventory_aversion, terminal_inventory_aversion, inventory_exponent, terminal_time)


class RunningInventoryPenalty(_InventoryAversion):
    """PnL - dt*phi*q'^p - alpha*[terminal step]*q'^p   (:116-146) -- BASELINE.json's "InventoryAdjustedPnL"."""
    KIND = _abi.MBT_REW_RUNNING_INVENTORY_PENALTY

    def __init__(self, per_step_inventory_aversion=0.01, terminal_inventory_aversion=0.0, inventory_exponent=2.0):
        super().__init__(per_step_inventory_aversion, terminal_inventory_aversion, inventory_exponent, None)

    def reset(self, initial_state):
        pass


CjCriterion = RunningInventoryPenalty  # the reference's alias (:141-143)
InventoryAdjustedPnL = RunningInventoryPenalty  # the name BASELINE.json's north_star uses for the same criterion


class ExponentialUtility(RewardFunction):
    """-exp(-risk_aversion * terminal wealth) on the terminal step, 0 otherwise   (:149-166)."""
    KIND = _abi.MBT_REW_EXP_UTILITY

    def __init__(self, risk_aversion=0.1):
        super().__init__()
        self.risk_aversion = risk_aversion

    def reset(self, initial_state):
        pass
