"""Closed-form baseline agents (reference: mbt_gym/agents/BaselineAgents.py).

Agents are CALLERS of the hot path: `get_action(obs)` runs on the host exactly as in the reference, so existing
rollout loops (`generate_trajectory`, SB3-style loops) work unchanged.  Each agent additionally offers
`to_policy(env)`: the same rule as an `mbt_policy` for the fused on-device rollout (`env.rollout_summary`), where the
per-step host round trip disappears.  Policy constants are formed with the same Python expressions in both places, so
the device reproduces the host agent bit-for-bit -- for environments with raw observations
(`normalise_observation_space=False`); the state-dependent policies refuse normalised ones (`_require_raw_observations`).
"""
import warnings
from copy import deepcopy

import numpy as np

from .. import _abi
from ..gym.index_names import ASK_INDEX, BID_INDEX, CASH_INDEX, INVENTORY_INDEX, TIME_INDEX, ASSET_PRICE_INDEX
from ..gym.ModelDynamics import LimitOrderModelDynamics, TradinghWithSpeedModelDynamics
from ..gym.TradingEnvironment import TradingEnvironment
from ..rewards.RewardFunctions import CjMmCriterion, PnL
from .Agent import Agent


def _require_raw_observations(env, agent):
    """The closed-form agents read inventory and time straight from the observation (`state[:, INVENTORY_INDEX]`,
    `state[:, TIME_INDEX]`, reference :64-65,117-119,205), i.e. they are written for
    `normalise_observation_space=False` (what the reference's notebooks construct).  The on-device policy reads the raw
    state, so with normalised observations it would NOT reproduce `get_action(env.step(...)[0])`: refuse instead."""
    if getattr(env, "normalise_observation_space_", False):
        raise ValueError(f"{type(agent).__name__}.to_policy needs an environment built with "
                         "normalise_observation_space=False: the agent's formula reads raw inventory and time")


def _fixed_policy(row):
    pol = _abi.mbt_policy()
    pol.kind = _abi.MBT_POL_FIXED
    for j, v in enumerate(np.asarray(row, float).reshape(-1)[: _abi.MBT_MAX_ACTION_DIM]):
        pol.fixed[j] = float(v)
    return pol


class RandomAgent(Agent):
    """One action-space sample per step, repeated for every trajectory (:15-22)."""

    def __init__(self, env, seed=None):
        self.action_space = deepcopy(env.action_space)
        self.action_space.seed(seed)
        self.num_trajectories = env.num_trajectories

    def get_action(self, state):
        return np.repeat(self.action_space.sample().reshape(1, -1), self.num_trajectories, axis=0)


class FixedActionAgent(Agent):
    def __init__(self, fixed_action, env):
        self.fixed_action, self.env = fixed_action, env

    def get_action(self, state):
        return np.repeat(np.asarray(self.fixed_action).reshape(1, -1), self.env.num_trajectories, axis=0)

    def to_policy(self, env=None):
        return _fixed_policy(self.fixed_action)


class FixedSpreadAgent(Agent):
    def __init__(self, env, half_spread=1.0, offset=0.0):
        self.half_spread, self.offset, self.env = half_spread, offset, env

    def _row(self):
        return np.array([[self.half_spread - self.offset, self.half_spread + self.offset]])

    def get_action(self, state):
        return np.repeat(self._row(), self.env.num_trajectories, axis=0)

    def to_policy(self, env=None):
        return _fixed_policy(self._row())


class HumanAgent(Agent):
    """Asks for the two half-spreads on the terminal at every step (:45-49); single-trajectory play only."""

    def get_action(self, state):
        prompt = "Current state is {}. How large do you want to set {} half spread? "
        return np.array([float(input(prompt.format(state, side))) for side in ("midprice-bid", "ask-midprice")])


class AvellanedaStoikovAgent(Agent):
    """Avellaneda & Stoikov (2008) quotes: reservation-price shift q*gamma*sigma^2*(T-t) around a spread
    gamma*sigma^2*(T-t) + (2/gamma)*log(1+gamma/kappa)   (:52-83)."""

    def __init__(self, risk_aversion=0.1, env=None):
        self.risk_aversion = risk_aversion
        self.env = env or TradingEnvironment()
        assert isinstance(self.env, TradingEnvironment)
        self.terminal_time = self.env.terminal_time
        self.volatility = self.env.model_dynamics.midprice_model.volatility
        self.rate_of_arrival = self.env.model_dynamics.arrival_model.intensity
        self.fill_exponent = self.env.model_dynamics.fill_probability_model.fill_exponent

    def _fill_component(self):
        if self.risk_aversion == 0:
            return 2 / self.fill_exponent  # limit gamma -> 0
        return 2 / self.risk_aversion * np.log(1 + self.risk_aversion / self.fill_exponent)

    def get_action(self, state):
        q, t = state[:, INVENTORY_INDEX], state[:, TIME_INDEX]
        tau = self.terminal_time - t
        adjustment = q * self.risk_aversion * self.volatility ** 2 * tau
        if self.risk_aversion == 0:
            spread = self._fill_component() * np.ones_like(tau)
        else:
            spread = self.risk_aversion * self.volatility ** 2 * tau + self._fill_component()
        action = np.stack([adjustment + spread / 2, -adjustment + spread / 2], axis=1)
        if action.min() < 0:
            warnings.warn("Avellaneda-Stoikov agent is quoting a negative spread")
        return action

    def to_policy(self, env=None):
        _require_raw_observations(env or self.env, self)
        pol = _abi.mbt_policy()
        pol.kind = _abi.MBT_POL_AVELLANEDA_STOIKOV
        pol.as_gamma = float(self.risk_aversion)
        pol.as_sigma_sq = float(self.volatility ** 2)
        pol.as_fill_comp = float(self._fill_component())
        pol.as_terminal_time = float(self.terminal_time)
        return pol


class CarteaJaimungalMmAgent(Agent):
    """Cartea, Jaimungal & Penalva (2015) ch. 10 market maker: depths from h(t,q) = log(omega)/kappa with
    omega(t) = expm(A (T-t)) z  (eq. 10.11)   (:86-170)."""

    def __init__(self, env=None):
        from scipy.linalg import expm  # noqa: F401  (fail early if scipy is missing)

        self.env = env or TradingEnvironment()
        assert isinstance(self.env.model_dynamics, LimitOrderModelDynamics), "Trader must be type LimitOrderTrader"
        assert isinstance(self.env.reward_function, (CjMmCriterion, PnL)), "Reward function for CjMmAgent is incorrect."
        self.kappa = self.env.model_dynamics.fill_probability_model.fill_exponent
        self.num_trajectories = self.env.num_trajectories
        self.inventory_neutral = isinstance(self.env.reward_function, PnL)
        if self.inventory_neutral:
            self.risk_neutral_action = 1 / self.kappa * np.ones((env.num_trajectories, env.action_space.shape[0]))
            return
        self.phi = env.reward_function.per_step_inventory_aversion
        self.alpha = env.reward_function.terminal_inventory_aversion
        assert self.env.reward_function.inventory_exponent == 2.0, "Inventory exponent must be = 2."
        self.terminal_time = self.env.terminal_time
        self.lambdas = self.env.model_dynamics.arrival_model.intensity
        self.max_inventory = env.max_inventory
        self.a_matrix, self.z_vector = self._calculate_a_and_z()
        self.large_depth = 10_000
        self._h_cache = {}

    # -- closed form
    def _calculate_a_and_z(self):
        Q = self.max_inventory
        q = Q - np.arange(2 * Q + 1)  # row i holds inventory Q - i
        A = np.diag(-self.phi * self.kappa * q.astype(float) ** 2)
        A += np.diag(np.full(2 * Q, self.lambdas[BID_INDEX] * np.exp(-1)), 1)
        A += np.diag(np.full(2 * Q, self.lambdas[ASK_INDEX] * np.exp(-1)), -1)
        z = np.exp(-self.alpha * self.kappa * q.astype(float) ** 2).reshape(-1, 1)
        return A, z

    def _calculate_omega(self, current_time):
        from scipy.linalg import expm

        return np.matmul(expm(self.a_matrix * (self.terminal_time - current_time)), self.z_vector)

    def _calculate_ht(self, current_time):
        key = float(current_time)
        if key not in self._h_cache:
            if len(self._h_cache) > 4096:
                self._h_cache.clear()
            self._h_cache[key] = 1 / self.kappa * np.log(self._calculate_omega(current_time))
        return self._h_cache[key]

    def _deltas_for_indices(self, h_t, indices):
        Q2 = 2 * self.max_inventory
        up, down = np.clip(indices + 1, 0, Q2), np.clip(indices - 1, 0, Q2)
        h0, h_up, h_down = h_t[indices], h_t[up], h_t[down]
        bid = 1 / self.kappa - h_up + h0 + self.large_depth * (h_up == h0)
        ask = 1 / self.kappa - h_down + h0 + self.large_depth * (h_down == h0)
        return bid.reshape(-1), ask.reshape(-1)

    def _calculate_deltas(self, current_time, inventories):
        h_t = self._calculate_ht(current_time)
        idx = np.clip(self.max_inventory + inventories, 0, 2 * self.max_inventory).astype(int)
        deltas = np.zeros((len(idx), 2))
        deltas[:, BID_INDEX], deltas[:, ASK_INDEX] = self._deltas_for_indices(h_t, idx)
        return deltas

    def get_action(self, state):
        if self.inventory_neutral:
            return self.risk_neutral_action
        assert state[0, TIME_INDEX] == state[-1, TIME_INDEX], \
            "CarteaJaimungalMmAgent needs to be called on a tensor with a uniform time stamp."
        return self._calculate_deltas(current_time=state[0, TIME_INDEX], inventories=state[:, INVENTORY_INDEX])

    def calculate_true_value_function(self, state):
        h_t = self._calculate_ht(state[0, TIME_INDEX])
        idx = np.clip(self.max_inventory + state[:, INVENTORY_INDEX], 0, 2 * self.max_inventory).astype(int)
        return h_t[idx] + state[:, CASH_INDEX] + state[:, INVENTORY_INDEX] * state[:, ASSET_PRICE_INDEX]

    # -- device form: (bid, ask) depth per (decision time, inventory index)
    def decision_times(self, env):
        """The clock values the remaining steps of the running episode will present to the agent."""
        t = float(env._native.clock()["time"]) if env._native is not None and env._started else float(env._get_start_time())
        times = []
        while not (t >= env.terminal_time - env.step_size / 2):
            times.append(t)
            t = t + env.step_size
        return times

    def to_policy(self, env=None):
        env = env or self.env
        if self.inventory_neutral:
            return _fixed_policy(self.risk_neutral_action[0])
        _require_raw_observations(env, self)
        times = self.decision_times(env)
        cols = 2 * self.max_inventory + 1
        table = np.empty((len(times), cols, 2))
        all_idx = np.arange(cols)
        for k, t in enumerate(times):
            table[k, :, 0], table[k, :, 1] = self._deltas_for_indices(self._calculate_ht(t), all_idx)
        pol = _abi.mbt_policy()
        pol.kind = _abi.MBT_POL_CJ_MM_TABLE
        pol.table_rows, pol.table_cols, pol.inv_offset = len(times), cols, int(self.max_inventory)
        self._table = np.ascontiguousarray(table)  # keep alive while the policy struct is in use
        pol.table = self._table.ctypes.data
        return pol


class CarteaJaimungalOeAgent(Agent):
    """Cartea, Jaimungal & Penalva (2015) p.147 optimal liquidation speed (:173-210): depends on time only."""

    def __init__(self, phi=2 * 10 ** (-4), alpha=0.0001, env=None):
        self.phi, self.alpha = phi, alpha
        self.env = env or TradingEnvironment()
        self.price_impact_model = env.model_dynamics.price_impact_model
        assert isinstance(self.env.model_dynamics, TradinghWithSpeedModelDynamics), \
            "Trader must be type TradinghWithSpeedTrader"
        self.terminal_time = self.env.terminal_time
        self.temporary_price_impact = self.price_impact_model.temporary_impact_coefficient
        self.permanent_price_impact = self.price_impact_model.permanent_impact_coefficient
        self.num_trajectories = self.env.num_trajectories

    def _speed(self, current_time):
        k, b = self.temporary_price_impact, self.permanent_price_impact
        gamma = np.sqrt(self.phi / k)
        zeta = (self.alpha - 0.5 * b + np.sqrt(k * self.phi)) / (self.alpha - 0.5 * b - np.sqrt(k * self.phi))
        q0 = self.env.initial_inventory
        left = self.terminal_time - current_time
        speed = gamma * q0 * ((zeta * np.exp(gamma * left) + np.exp(-gamma * left))
                              / (zeta * np.exp(gamma * self.terminal_time) - np.exp(-gamma * self.terminal_time)))
        return -np.sign(q0) * speed

    def get_action(self, state):
        action = np.zeros((self.num_trajectories, 1))
        action[:, :] = self._speed(state[0, TIME_INDEX])
        return action

    def to_policy(self, env=None):
        env = env or self.env
        _require_raw_observations(env, self)
        times = CarteaJaimungalMmAgent.decision_times(self, env)
        self._table = np.ascontiguousarray([[self._speed(t)] for t in times], dtype=float)
        pol = _abi.mbt_policy()
        pol.kind = _abi.MBT_POL_SCHEDULE
        pol.table_rows = len(times)
        pol.table = self._table.ctypes.data
        return pol
