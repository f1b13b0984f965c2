#!/usr/bin/env python
"""Per-step time of the step kernel for the configurations a user of the reference is most likely to run, device-resident,
150 back-to-back steps between two CUDA events at N = 2^20 (float64): which kernel variant each one selects and what it
costs.  `python tools/variant_step_times.py`"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mbt_gym_b200.gym import ModelDynamics as MD  # noqa: E402
from mbt_gym_b200.gym.TradingEnvironment import TradingEnvironment  # noqa: E402
from mbt_gym_b200.rewards.RewardFunctions import CjMmCriterion, RunningInventoryPenalty  # noqa: E402
from mbt_gym_b200.stochastic_processes import arrival_models as AM, fill_probability_models as FM, midprice_models as MM  # noqa: E402

N = 1 << 20
DT = 1.0 / 200


def market(mid="bm", dyn="limit", fill="exp", arr="poisson"):
    kw = dict(step_size=DT, num_trajectories=N)
    m = {"bm": lambda: MM.BrownianMotionMidpriceModel(volatility=2.0, **kw),
         "gbm": lambda: MM.GeometricBrownianMotionMidpriceModel(drift=0.05, volatility=0.2, **kw),
         "bm_jump": lambda: MM.BrownianMotionJumpMidpriceModel(volatility=2.0, jump_size=0.1, **kw),
         "ou_jump": lambda: MM.OuJumpMidpriceModel(mean_reversion_level=100.0, mean_reversion_speed=0.1, volatility=2.0, jump_size=0.1, **kw)}[mid]()
    a = (AM.PoissonArrivalModel(**kw) if arr == "poisson" else AM.PoissonArrivalNonLinearModel(**kw))
    f = {"exp": lambda: FM.ExponentialFillFunction(**kw), "power": lambda: FM.PowerFillFunction(**kw),
         "triangular": lambda: FM.TriangularFillFunction(**kw)}[fill]()
    if dyn == "limit":
        return MD.LimitOrderModelDynamics(midprice_model=m, arrival_model=a, fill_probability_model=f, num_trajectories=N)
    if dyn == "touch":
        return MD.AtTheTouchModelDynamics(midprice_model=m, arrival_model=a, num_trajectories=N)
    return MD.LimitAndMarketOrderModelDynamics(midprice_model=m, arrival_model=a, fill_probability_model=f, num_trajectories=N)


def run(name, **kw):
    env = TradingEnvironment(num_trajectories=N, seed=3, **kw)
    env.reset_device()
    native = env._native
    A = native.A
    info = native.kernel_info()
    acts = [torch.full((N, A), 0.3, dtype=torch.float64, device="cuda") for _ in range(8)]
    obs = [torch.empty((N, native.Dout), dtype=torch.float64, device="cuda") for _ in range(8)]
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
    kern = (f"run-time specialised, {info['step_registers']} regs" if info["step_is_jit"] else f"ahead-of-time variant {info['aot_variant']}")
    print(f"{name:72s} {1e3 * a.elapsed_time(b) / 150:7.2f} us per step   [{kern}]")
    env.close()


if __name__ == "__main__":
    plain = dict(normalise_action_space=False, normalise_observation_space=False)
    run("default constructor: normalised actions + observations, PnL")
    run("  + CjMmCriterion", reward_function=CjMmCriterion(0.01, 0.001, 2.0, 1.0), max_inventory=100)
    run("  + RunningInventoryPenalty", reward_function=RunningInventoryPenalty(0.01, 0.001, 2.0))
    run("normalised observations only (runtime-flag variant)", normalise_action_space=False)
    run("plain spaces, PnL", **plain)
    run("plain spaces, RunningInventoryPenalty", reward_function=RunningInventoryPenalty(0.01, 0.001, 2.0), **plain)
    # configurations that ran the generic kernel in round 1 (29 us, 64 registers)
    run("GBM midprice, plain", model_dynamics=market(mid="gbm"), **plain)
    run("GBM midprice + non-linear Poisson arrivals, default spaces", model_dynamics=market(mid="gbm", arr="nonlinear"))
    run("BM-jump midprice, plain", model_dynamics=market(mid="bm_jump"), **plain)
    run("OU-jump midprice, RunningInventoryPenalty, default spaces", model_dynamics=market(mid="ou_jump"),
        reward_function=RunningInventoryPenalty(0.01, 0.001, 2.0))
    run("at-the-touch dynamics, plain", model_dynamics=market(dyn="touch"), **plain)
    run("limit + market orders, plain", model_dynamics=market(dyn="limit_and_market"), **plain)
    run("power fill function (batch reduction + step), plain", model_dynamics=market(fill="power"), **plain)
    run("triangular fill function (batch reduction + step), default spaces", model_dynamics=market(fill="triangular"))
