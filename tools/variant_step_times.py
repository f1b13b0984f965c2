#!/usr/bin/env python
"""Per-step time of the step kernel for the configurations a user of the reference is most likely to run, device-resident,
150 back-to-back steps between two CUDA events at N = 2^20 (float64): which kernel variant each one selects and what it
costs.  `python tools/variant_step_times.py`"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mbt_gym_b200.gym.TradingEnvironment import TradingEnvironment  # noqa: E402
from mbt_gym_b200.rewards.RewardFunctions import CjMmCriterion, RunningInventoryPenalty  # noqa: E402

N = 1 << 20


def run(name, **kw):
    env = TradingEnvironment(num_trajectories=N, seed=3, **kw)
    env.reset_device()
    native = env._native
    acts = [torch.zeros((N, 2), dtype=torch.float64, device="cuda") for _ in range(8)]
    obs = [torch.empty((N, 4), dtype=torch.float64, device="cuda") for _ in range(8)]
    rew = [torch.empty((N,), dtype=torch.float64, device="cuda") for _ in range(8)]
    for k in range(10):
        native.step(acts[k % 8], obs[k % 8], rew[k % 8], mem=1)
    env.reset_device()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(150):
        native.step(acts[k % 8], obs[k % 8], rew[k % 8], mem=1)
    b.record()
    torch.cuda.synchronize()
    print(f"{name:72s} {1e3 * a.elapsed_time(b) / 150:7.2f} us per step")
    env.close()


if __name__ == "__main__":
    plain = dict(normalise_action_space=False, normalise_observation_space=False)
    run("default constructor: normalised actions + observations, PnL")
    run("  + CjMmCriterion", reward_function=CjMmCriterion(0.01, 0.001, 2.0, 1.0), max_inventory=100)
    run("  + RunningInventoryPenalty", reward_function=RunningInventoryPenalty(0.01, 0.001, 2.0))
    run("normalised observations only (runtime-flag variant)", normalise_action_space=False)
    run("plain spaces, PnL", **plain)
    run("plain spaces, RunningInventoryPenalty", reward_function=RunningInventoryPenalty(0.01, 0.001, 2.0), **plain)
