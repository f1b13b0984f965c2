#!/usr/bin/env python
"""Measure the host<->device copy rates that bound the host-buffer (e2e) path: pinned D2H, pinned H2D, both at once."""
import time

import torch

N = 1 << 20
dev = torch.device("cuda", 0)
d_obs = torch.empty((N, 5), dtype=torch.float64, device=dev)          # obs + rewards: 41.9 MB
d_act = torch.empty((N, 2), dtype=torch.float64, device=dev)          # actions: 16.8 MB
h_obs = torch.empty((N, 5), dtype=torch.float64).pin_memory()
h_act = torch.empty((N, 2), dtype=torch.float64).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=50):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def d2h():
    with torch.cuda.stream(s1):
        h_obs.copy_(d_obs, non_blocking=True)


def h2d():
    with torch.cuda.stream(s2):
        d_act.copy_(h_act, non_blocking=True)


def both():
    d2h(); h2d()


t = timed(d2h); print(f"D2H 41.9 MB alone: {t*1e3:.3f} ms  {41.943/t/1e3:.1f} GB/s")
t = timed(h2d); print(f"H2D 16.8 MB alone: {t*1e3:.3f} ms  {16.777/t/1e3:.1f} GB/s")
t = timed(both); print(f"both concurrently: {t*1e3:.3f} ms")
