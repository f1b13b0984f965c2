"""Adapter from a trained stable-baselines3 model to the `Agent` interface (reference: mbt_gym/agents/SbAgent.py:8-26).

A CALLER of the hot path.  stable-baselines3 is not imported here: anything with `predict(obs, deterministic=True)`,
`action_space` and (for `train`) `learn(total_timesteps=...)` works, which is all the reference uses of `BaseAlgorithm`.
"""
import numpy as np

from .Agent import Agent


class SbAgent(Agent):
    def __init__(self, model, reduced_training_indices=None, num_trajectories=None):
        self.model = model
        self.num_trajectories = num_trajectories or self.model.env.num_trajectories
        self.num_actions = self.model.action_space.shape[0]
        self.reduced_training = reduced_training_indices is not None
        if self.reduced_training:
            self.reduced_training_indices = reduced_training_indices

    def get_action(self, state):
        """Deterministic policy action for every row of `state`; the model sees only the columns it was trained on."""
        obs = state[:, self.reduced_training_indices] if self.reduced_training else state
        action, _ = self.model.predict(obs, deterministic=True)
        return np.asarray(action).reshape(obs.shape[0], self.num_actions)

    def train(self, total_timesteps=100000):
        self.model.learn(total_timesteps=total_timesteps)
