#!/usr/bin/env python
"""Zero-copy RL interaction loop: a torch policy network on the GPU acts on CUDA observation tensors and
`TradingEnvironment.step` consumes its CUDA action tensor directly -- nothing crosses PCIe.

The loop is what an on-device learner (e.g. an SB3-style PPO with device buffers) does per environment step:
    obs (N, D) cuda  ->  policy MLP  ->  action (N, A) cuda  ->  env.step  ->  obs, rewards (cuda), dones

    python examples/torch_policy_loop.py [--n 1048576] [--episodes 3]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from mbt_gym_b200.gym.TradingEnvironment import TradingEnvironment  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 20)
    ap.add_argument("--episodes", type=int, default=3)
    ap.add_argument("--precision", default="float32", choices=["float32", "float64"])
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    tdt = torch.float32 if args.precision == "float32" else torch.float64
    # the reference's default market (AS-like), normalised observation and action spaces in [-1, 1] as SB3 expects
    env = TradingEnvironment(num_trajectories=args.n, seed=7, precision=args.precision, device=0)
    policy = torch.nn.Sequential(torch.nn.Linear(4, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                                 torch.nn.Linear(64, 2), torch.nn.Tanh()).to(dev, tdt)
    total_steps, t_total = 0, 0.0
    for ep in range(args.episodes):
        obs = torch.from_numpy(env.reset()).to(dev)          # first observation of the episode comes through the host API
        ep_return = torch.zeros(args.n, dtype=tdt, device=dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        done = False
        with torch.no_grad():
            while not done:
                action = policy(obs)
                obs, rew, dones, _ = env.step(action)          # CUDA tensors in, CUDA tensors out
                ep_return += rew
                done = bool(dones[0])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if ep > 0:  # first episode warms up cuBLAS / the allocator
            total_steps += args.n * env.n_steps
            t_total += dt
        print(f"episode {ep}: mean return {ep_return.mean().item():8.3f}  {args.n * env.n_steps / dt:.3e} env-steps/s "
              f"({1e6 * dt / env.n_steps:.1f} us per step incl. the policy MLP)")
    if t_total:
        print(f"steady state: {total_steps / t_total:.3e} env-steps/s with a 4-64-64-2 tanh MLP policy ({args.precision})")
    env.close()


if __name__ == "__main__":
    main()
