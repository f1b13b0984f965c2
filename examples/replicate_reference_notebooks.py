#!/usr/bin/env python
"""Replicate the reference's two validation notebooks on the GPU with the fused on-device rollout.

  Test_1 - replicate_AS_original_results.ipynb : Avellaneda-Stoikov (2008) results tables, gamma = 0.1 and 0.01
      reference (N = 1000, seed 50):  gamma 0.1 : spread 1.49177  PnL 64.872139 +- 6.692567  q_T 0.201 +- 2.893544
                                      gamma 0.01: spread 1.349009 PnL 68.754417 +- 8.720076  q_T 0.230 +- 5.095989
  Test_2 - replicate_CJP_2015 ... value_function: sample mean of the total CjMm reward vs the closed-form value function
      reference value functions: 68.25583476, 73.22586344, 18.21929052, 36.32607427

    python examples/replicate_reference_notebooks.py [--n 1048576]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mbt_gym_b200.agents.BaselineAgents import AvellanedaStoikovAgent, CarteaJaimungalMmAgent  # noqa: E402
from mbt_gym_b200.gym.helpers.generate_trajectory import generate_results_table_fused  # noqa: E402
from mbt_gym_b200.gym.ModelDynamics import LimitOrderModelDynamics  # noqa: E402
from mbt_gym_b200.gym.TradingEnvironment import TradingEnvironment  # noqa: E402
from mbt_gym_b200.rewards.RewardFunctions import CjMmCriterion, PnL  # noqa: E402
from mbt_gym_b200.stochastic_processes.arrival_models import PoissonArrivalModel  # noqa: E402
from mbt_gym_b200.stochastic_processes.fill_probability_models import ExponentialFillFunction  # noqa: E402
from mbt_gym_b200.stochastic_processes.midprice_models import BrownianMotionMidpriceModel  # noqa: E402


def market_env(N, S0, sigma, lam, kappa, T, n_steps, reward, max_inventory, seed):
    dt = T / n_steps
    dyn = LimitOrderModelDynamics(
        midprice_model=BrownianMotionMidpriceModel(volatility=sigma, initial_price=S0, terminal_time=T, step_size=dt, num_trajectories=N),
        arrival_model=PoissonArrivalModel(intensity=np.array([lam, lam]), step_size=dt, num_trajectories=N),
        fill_probability_model=ExponentialFillFunction(fill_exponent=kappa, step_size=dt, num_trajectories=N),
        num_trajectories=N)
    return TradingEnvironment(terminal_time=T, n_steps=n_steps, reward_function=reward, model_dynamics=dyn,
                              max_inventory=max_inventory, seed=seed, num_trajectories=N, normalise_action_space=False,
                              normalise_observation_space=False)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 20)
    args = ap.parse_args()
    N = args.n

    print(f"== Test_1: Avellaneda-Stoikov tables, N = {N} trajectories (reference: N = 1000)")
    ref = {0.1: (1.49177, 64.872139, 6.692567, 0.201, 2.893544), 0.01: (1.349009, 68.754417, 8.720076, 0.23, 5.095989)}
    for gamma in (0.1, 0.01):
        env = market_env(N, 100.0, 2.0, 140.0, 1.5, 1.0, 200, PnL(), 200, seed=50)
        table, _ = generate_results_table_fused(env, AvellanedaStoikovAgent(risk_aversion=gamma, env=env))
        got = [table[k] for k in ("Mean spread", "Mean PnL", "Std PnL", "Mean terminal inventory", "Std terminal inventory")]
        se = ref[gamma][2] / np.sqrt(1000)
        print(f"  gamma={gamma:<5} GPU       spread {got[0]:.5f}  PnL {got[1]:.4f} +- {got[2]:.4f}  q_T {got[3]:+.4f} +- {got[4]:.4f}")
        print(f"              reference spread {ref[gamma][0]:.5f}  PnL {ref[gamma][1]:.4f} +- {ref[gamma][2]:.4f}  q_T {ref[gamma][3]:+.4f} +- {ref[gamma][4]:.4f}"
              f"   |dPnL| = {abs(got[1]-ref[gamma][1]):.3f} = {abs(got[1]-ref[gamma][1])/se:.2f} SE of the reference's sample")
        env.close()

    print(f"== Test_2: CJP-2015 closed-form value function vs simulated mean total reward, N = {N}")
    sets = [("set 0", 100.0, 2.0, 140.0, 1.5, 1.0, 1000, 68.25583476, 68.24261382, 12.23134896),
            ("set I", 150.0, 1.0, 100.0, 1.0, 1.0, 1000, 73.22586344, 72.98439169, 9.83044035),
            ("set II", 50.0, 1.5, 50.0, 2.0, 1.0, 2000, 18.21929052, 18.24885364, 6.06485497),
            ("set III", 50.0, 1.5, 50.0, 2.0, 2.0, 2000, 36.32607427, 36.45658480, 8.89859370)]
    for name, S0, sigma, lam, kappa, T, n_steps, ref_v, ref_mean, ref_std in sets:
        env = market_env(N, S0, sigma, lam, kappa, T, n_steps, CjMmCriterion(0.01, 0.001, 2.0, T), 100, seed=410)
        agent = CarteaJaimungalMmAgent(env=env)
        v = float(np.asarray(agent.calculate_true_value_function(np.array([[0.0, 0.0, 0.0, S0]] * 2))).reshape(-1)[0])
        env.reset()
        summary, returns, _q = env.rollout_summary(agent.to_policy(env), return_trajectory_stats=True)
        mean, std = returns.mean(), returns.std()
        print(f"  {name:8s} value fn {v:.8f} (reference {ref_v:.8f})   GPU mean {mean:.5f} +- {std:.5f} (SE {std/np.sqrt(N):.5f})"
              f"   reference sample {ref_mean:.5f} +- {ref_std:.5f} (SE {ref_std/np.sqrt(1000):.3f})   mean - value fn = {mean - v:+.5f}")
        env.close()


if __name__ == "__main__":
    main()
