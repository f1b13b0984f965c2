#!/usr/bin/env python
"""Launch each hot-path kernel once or twice at BASELINE size (2^20 trajectories) so `ncu` can capture it.

    ncu --set full --clock-control none --import-source on -k regex:'mbt_(rollout|fill_batch|step)' \
        -o gpurun_out/targets python tools/profile_targets.py [f64|f32]

Order of launches: reset, 2 x step (AS), rollout (fixed action), reset, rollout (Avellaneda-Stoikov policy),
then the power-fill market: reset, 2 x (fill_batch + step); then the reference's default-constructor flags (normalised
actions + observations, ahead-of-time variant 10): 2 x step; then a market only the run-time specialiser serves
(GBM midprice, non-linear Poisson arrivals, RunningInventoryPenalty, normalised observations): 2 x step.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mbt_gym_b200 import _abi, _lib  # noqa: E402


def main():
    prec = _abi.MBT_F32 if (len(sys.argv) > 1 and sys.argv[1] == "f32") else _abi.MBT_F64
    dt = np.float32 if prec == _abi.MBT_F32 else np.float64
    N, n_steps = 1 << 20, 200
    base = dict(precision=prec, num_trajectories=N, n_steps=n_steps, dynamics=_abi.MBT_DYN_LIMIT, midprice=_abi.MBT_MID_BM,
                arrival=_abi.MBT_ARR_POISSON, reward=_abi.MBT_REW_PNL, terminal_time=1.0, step_size=1.0 / n_steps,
                max_inventory=200, max_cash=200 * 108.0, mid_initial=100.0, mid_vol=2.0, mid_step=1.0 / n_steps,
                arr_rate=[140.0, 140.0], arr_step=1.0 / n_steps, rew_terminal_time=1.0)
    import torch

    tdt = torch.float32 if prec == _abi.MBT_F32 else torch.float64
    act = torch.full((N, 2), 0.7, dtype=tdt, device="cuda")
    obs = torch.empty((N, 4), dtype=tdt, device="cuda")
    rew = torch.empty((N,), dtype=tdt, device="cuda")
    torch.cuda.synchronize()

    env = _lib.NativeEnv(_abi.new_config(fill=_abi.MBT_FILL_EXPONENTIAL, fill_exponent=1.5, **base), device=0)
    env.seed(50)
    env.reset()
    for _ in range(2):
        env.step(act, obs, rew, mem=_abi.MBT_MEM_DEVICE)
    env.sync()
    pol = _abi.mbt_policy()
    pol.kind = _abi.MBT_POL_FIXED
    pol.fixed[0] = pol.fixed[1] = 0.7
    ret = np.empty(N, dt)
    s = env.rollout(pol, ret)
    print("fixed-action rollout: steps", s.steps, "mean return", s.sum_return / N)
    env.reset()
    pol.kind = _abi.MBT_POL_AVELLANEDA_STOIKOV
    pol.as_gamma, pol.as_sigma_sq = 0.1, 4.0
    pol.as_fill_comp = 2 / 0.1 * np.log(1 + 0.1 / 1.5)
    pol.as_terminal_time = 1.0
    s = env.rollout(pol, ret)
    print("Avellaneda-Stoikov rollout: steps", s.steps, "mean return", s.sum_return / N)
    env.close()

    env = _lib.NativeEnv(_abi.new_config(fill=_abi.MBT_FILL_POWER, fill_exponent=1.5, fill_multiplier=1.5, **base), device=0)
    env.seed(50)
    env.reset()
    for _ in range(2):
        env.step(act, obs, rew, mem=_abi.MBT_MEM_DEVICE)
    env.sync()
    print("power-fill steps done; mean reward", float(rew.mean()))
    env.close()

    norm = dict(normalise_action=1, normalise_obs=1, act_low=[0.0, 0.0, 0, 0], act_grad=[1.535, 1.535, 0, 0],
                obs_low=[-21600.0, -200.0, 0.0, 92.0, 0, 0, 0, 0], obs_grad=[21600.0, 200.0, 0.5, 8.0, 1, 1, 1, 1])
    env = _lib.NativeEnv(_abi.new_config(fill=_abi.MBT_FILL_EXPONENTIAL, fill_exponent=1.5, **base, **norm), device=0)
    env.seed(50)
    env.reset()
    for _ in range(2):
        env.step(act, obs, rew, mem=_abi.MBT_MEM_DEVICE)
    env.sync()
    print("normalised AS steps done;", env.kernel_info())
    env.close()

    jit = dict(base, midprice=_abi.MBT_MID_GBM, arrival=_abi.MBT_ARR_POISSON_NONLINEAR,
               reward=_abi.MBT_REW_RUNNING_INVENTORY_PENALTY, mid_drift=0.05, mid_vol=0.2)
    norm_obs_only = dict(norm, normalise_action=0)
    env = _lib.NativeEnv(_abi.new_config(fill=_abi.MBT_FILL_EXPONENTIAL, fill_exponent=1.5, rew_phi=0.01, rew_alpha=0.001,
                                         **jit, **norm_obs_only), device=0)
    env.seed(50)
    env.reset()
    for _ in range(2):
        env.step(act, obs, rew, mem=_abi.MBT_MEM_DEVICE)
    env.sync()
    print("run-time specialised GBM market steps done;", env.kernel_info())
    env.close()


if __name__ == "__main__":
    main()
