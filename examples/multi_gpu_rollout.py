#!/usr/bin/env python
"""BASELINE configs[4] as a user would write it: optimal execution with an OU midprice, the trajectories sharded over the
GPUs of one box (one process per GPU), episodes run by the fused on-device rollout, and the ONE exchange of the path -- the
episode summary (all-reduce) and the per-trajectory returns (all-gather) -- done by the library over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/multi_gpu_rollout.py
    python examples/multi_gpu_rollout.py                       # one GPU: same code, no group

Replaces the reference's process fan-out (mbt_gym/gym/MultiprocessTradingEnv.py:72-116)."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mbt_gym_b200 import sharding  # noqa: E402
from mbt_gym_b200.agents.BaselineAgents import CarteaJaimungalOeAgent  # noqa: E402
from mbt_gym_b200.gym.ModelDynamics import TradinghWithSpeedModelDynamics  # noqa: E402
from mbt_gym_b200.gym.TradingEnvironment import TradingEnvironment  # noqa: E402
from mbt_gym_b200.rewards.RewardFunctions import CjOeCriterion  # noqa: E402
from mbt_gym_b200.stochastic_processes.midprice_models import OuMidpriceModel  # noqa: E402
from mbt_gym_b200.stochastic_processes.price_impact_models import TemporaryAndPermanentPriceImpact  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trajectories", dest="n", type=int, default=8 * (1 << 20), help="trajectories over ALL GPUs")
    ap.add_argument("--episodes", type=int, default=3)
    args = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lo, hi = sharding.shard_bounds(args.n, world, rank)          # contiguous shard of global trajectory ids
    n, T, n_steps = hi - lo, 1.0, 200
    dt = T / n_steps
    dyn = TradinghWithSpeedModelDynamics(
        midprice_model=OuMidpriceModel(mean_reversion_level=100.0, mean_reversion_speed=1.0, volatility=2.0, initial_price=100.0,
                                       terminal_time=T, step_size=dt, num_trajectories=n),
        price_impact_model=TemporaryAndPermanentPriceImpact(0.01, 0.01, n_steps=n_steps, terminal_time=T, num_trajectories=n),
        num_trajectories=n)
    env = TradingEnvironment(terminal_time=T, n_steps=n_steps, reward_function=CjOeCriterion(0.01, 0.001, 2.0, T),
                             model_dynamics=dyn, initial_inventory=100, max_inventory=10_000, seed=1234, num_trajectories=n,
                             normalise_action_space=False, normalise_observation_space=False, device=local, traj_offset=lo)
    native = env._ensure_native()
    if world > 1:
        sharding.create_group(native)                             # ncclCommInitRank inside libmbt_b200
    agent = CarteaJaimungalOeAgent(phi=0.01, alpha=0.001, env=env)  # closed-form liquidation speed: a time schedule
    tdt = torch.float64
    returns = [torch.empty(n, dtype=tdt, device="cuda") for _ in range(2)]          # two buffers: the gather of episode k
    gathered = [torch.empty(args.n, dtype=tdt, device="cuda") for _ in range(2)]    # overlaps the rollout of episode k+1
    for ep in range(args.episodes):
        env.reset_device()
        policy = agent.to_policy(env)
        if world > 1:
            summary = native.group_rollout(policy, returns[ep % 2], gathered[ep % 2])   # summary of ALL trajectories
        else:
            summary = native.rollout(policy, returns[ep % 2], None, mem=1)
            gathered[ep % 2] = returns[ep % 2]
        table = sharding.results_table(summary, action_dim=1)
        if rank == 0:
            print(f"episode {ep}: {summary.count} trajectories x {summary.steps} steps   mean return {table['Mean PnL']:.4f} "
                  f"+- {table['Std PnL']:.4f}   mean terminal inventory {table['Mean terminal inventory']:.4f}")
    if world > 1:
        native.group_wait()                                       # the last gather
    inv, counts, (below, above) = env.inventory_histogram(lo=0, hi=100, group_sum=world > 1)
    if rank == 0:
        g = gathered[(args.episodes - 1) % 2]
        print(f"gathered returns: {g.numel()} values, mean {float(g.mean()):.4f} (global-id order, identical on every rank)")
        print("terminal inventory histogram (all GPUs):", {int(q): int(c) for q, c in zip(inv, counts) if c}, "outside:", below + above)
    env.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
