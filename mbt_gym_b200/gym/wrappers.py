"""Observation / reward wrappers around the environment surface (reference: mbt_gym/gym/wrappers.py).
Pure host-side glue on top of `TradingEnvironment` -- they work on any object exposing reset/step/spaces."""
import numpy as np

from ..spaces import Box
from .index_names import INVENTORY_INDEX, TIME_INDEX


class _Wrapper:
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)


class ReduceStateSizeWrapper(_Wrapper):
    """Keep only the listed observation columns (default: inventory and time)  (:10-43).

    On a `TradingEnvironment` with increasing column indices the selection is FUSED into the kernel's observation
    store (`env.select_observation_columns`): no (N, D) -> (N, k) fancy-index copy on the host and k/D of the D2H bytes.
    Anything else (another wrapper underneath, a permutation of columns) falls back to host-side indexing."""

    def __init__(self, env, list_of_state_indices=(INVENTORY_INDEX, TIME_INDEX), fuse=True):
        super().__init__(env)
        self.list_of_state_indices = list(list_of_state_indices)
        space = env.observation_space
        self.observation_space = Box(low=space.low[self.list_of_state_indices],
                                     high=space.high[self.list_of_state_indices], dtype=np.float64)
        idx = self.list_of_state_indices
        self._fused = bool(fuse and hasattr(env, "select_observation_columns") and idx == sorted(set(idx))
                           and getattr(env, "_obs_columns", None) is None)
        if self._fused:
            env.select_observation_columns(idx)

    def observation(self, observation):
        return observation if self._fused else observation[:, self.list_of_state_indices]

    def reset(self):
        return self.observation(self.env.reset())

    def reset_device(self):
        """First observation as a CUDA tensor (see `TradingEnvironment.reset_device`), columns selected."""
        return self.observation(self.env.reset_device())

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        return self.observation(obs), reward, done, info

    @property
    def spec(self):
        return getattr(self.env, "spec", None)


class NormaliseASObservation(_Wrapper):
    """Affine map of the observation box to [-1, 1]  (:46-76).  NB the reference's `step` returns
    `obs / normalisation_factor` (not the same map as its `reset`); that behaviour is kept."""

    def __init__(self, env):
        super().__init__(env)
        self.normalisation_factor = 2 / (env.observation_space.high - env.observation_space.low)
        self.normalisation_offset = (env.observation_space.high + env.observation_space.low) / 2
        self.observation_space = Box(low=-np.ones(env.observation_space.shape), high=np.ones(env.observation_space.shape),
                                     dtype=np.float64)

    def reset(self):
        return (self.env.reset() - self.normalisation_offset) * self.normalisation_factor

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        return obs / self.normalisation_factor, reward, done, info


class RemoveTerminalRewards(_Wrapper):
    """Scale the terminal step's reward by phi / alpha of the reward function  (:79-105).  (The reference tests
    `if done:` on the dones array, which only works for one trajectory; here the uniform flag `done[0]` is used.)"""

    def __init__(self, env, num_final_steps=5):
        super().__init__(env)

    def reset(self):
        return self.env.reset()

    def step(self, action):
        state, reward, done, _ = self.env.step(action)
        if np.asarray(done).reshape(-1)[0]:
            rf = self.env.reward_function
            reward = reward * (rf.per_step_inventory_aversion / rf.terminal_inventory_aversion)
        return state, reward, done, {}
