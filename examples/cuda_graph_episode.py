#!/usr/bin/env python
"""A whole episode -- reset, policy network, env step, 200 times -- as ONE CUDA graph.

Small batches are launch-bound: a 4 096-trajectory step is a few microseconds of GPU work behind ~10 launches of
Python / ctypes / torch dispatch.  Capturing the episode into a CUDA graph removes the host from the loop; the
device-resident counter base (`env.fold_counters()`, C ABI `mbt_fold_counters`) makes every replay a NEW episode with
fresh random numbers, bit-identical to stepping the same episodes eagerly.

    python examples/cuda_graph_episode.py [--n 4096] [--episodes 20] [--precision float32]
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
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--episodes", type=int, default=20)
    ap.add_argument("--precision", default="float32", choices=["float32", "float64"])
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    tdt = torch.float32 if args.precision == "float32" else torch.float64
    torch.manual_seed(0)
    policy = torch.nn.Sequential(torch.nn.Linear(4, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh(),
                                 torch.nn.Linear(64, 2), torch.nn.Tanh()).to(dev, tdt)

    def make_env():
        return TradingEnvironment(num_trajectories=args.n, seed=7, precision=args.precision, device=0)

    # ---- eager: one Python iteration per env-step
    env = make_env()
    n_steps = env.n_steps
    eager_returns = []
    with torch.no_grad():
        for ep in range(args.episodes + 1):
            if ep == 1:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            obs = env.reset_device()
            ret = torch.zeros(args.n, dtype=tdt, device=dev)
            for _ in range(n_steps):
                obs, rew, _dones, _ = env.step(policy(obs))
                ret += rew
            eager_returns.append(ret.mean().item())
    torch.cuda.synchronize()
    t_eager = (time.perf_counter() - t0) / args.episodes
    env.close()

    # ---- one CUDA graph per episode
    env = make_env()
    s = torch.cuda.Stream()
    obs_t = torch.zeros((args.n, 4), dtype=tdt, device=dev)
    ret_t = torch.zeros(args.n, dtype=tdt, device=dev)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        env._ensure_native().set_stream(s.cuda_stream)  # bind the handle to the capture stream
        for _ in range(3):
            policy(obs_t)                                # warm up cuBLAS on the side stream (torch's capture recipe)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    env.prepare_capture()                                # needed if the env was already stepped (harmless on a fresh one)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s), torch.no_grad():
        env.reset_device(out=obs_t)
        ret_t.zero_()
        for _ in range(n_steps):
            o, rew, _dones, _ = env.step(policy(obs_t))
            obs_t.copy_(o)
            ret_t += rew
        env.fold_counters()                              # LAST: replays draw fresh random numbers
    graph_returns = []
    for ep in range(args.episodes + 1):
        if ep == 1:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        graph.replay()
        if ep < 3:
            graph_returns.append(ret_t.mean().item())    # (synchronises; only for the equality check below)
    torch.cuda.synchronize()
    t_graph = (time.perf_counter() - t0) / args.episodes
    env.close()

    same = all(abs(a - b) <= 1e-6 * max(1.0, abs(a)) for a, b in zip(eager_returns[:3], graph_returns))
    print(f"N = {args.n}, {n_steps} steps per episode, 4-64-64-2 tanh MLP policy, {args.precision}")
    print(f"  eager loop : {1e6 * t_eager / n_steps:8.1f} us per step  {args.n * n_steps / t_eager:.3e} env-steps/s")
    print(f"  CUDA graph : {1e6 * t_graph / n_steps:8.1f} us per step  {args.n * n_steps / t_graph:.3e} env-steps/s"
          f"   ({t_eager / t_graph:.1f}x)")
    print(f"  first episodes' mean returns, eager vs graph: {eager_returns[:3]} vs {graph_returns}  same: {same}")


if __name__ == "__main__":
    main()
