#!/usr/bin/env python
"""Does the L2 access-policy window of one handle slow down OTHER work on the device?  Step time at 2^22 / 2^24 trajectories
(no window: the state does not fit) measured alone and after a 2^20-trajectory handle (which sets the window and raises
the persisting-L2 set-aside) has run in the same process; plus a plain torch copy before / after.
    python tools/l2_window_probe.py            (MBT_L2_PERSIST=0 for the control run)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from mbt_gym_b200 import _abi  # noqa: E402


def step_us(n, stream, steps=30):
    f = bench.make_env("as", "float64", n, 0, 0)
    e = f._ensure_native()
    e.set_stream(stream.cuda_stream)
    b = bench.device_buffers(torch, n, e.A, e.D, torch.float64, 8, 0.7)
    acts, obs, rew, n_sets, _ = b
    e.reset(obs[0], mem=_abi.MBT_MEM_DEVICE)
    for k in range(10):
        e.step(acts[k % n_sets], obs[k % n_sets], rew[k % n_sets], mem=_abi.MBT_MEM_DEVICE)
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for k in range(steps):
        e.step(acts[k % n_sets], obs[k % n_sets], rew[k % n_sets], mem=_abi.MBT_MEM_DEVICE)
    z.record(stream)
    torch.cuda.synchronize()
    e.set_stream(None)
    f.close()
    return 1e3 * a.elapsed_time(z) / steps


def copy_gbs(stream):
    x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    y = torch.empty_like(x)
    y.copy_(x)
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(5):
        y.copy_(x)
    z.record(stream)
    torch.cuda.synchronize()
    return 5 * 2 * (1 << 30) / (a.elapsed_time(z) * 1e-3) / 1e9


stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
print("L2 window:", os.environ.get("MBT_L2_PERSIST", "default (on)"))
print("copy before any handle: %.0f GB/s" % copy_gbs(stream))
for n in (1 << 24, 1 << 22):
    print("N = 2^%d alone: %.1f us/step" % (n.bit_length() - 1, step_us(n, stream)))
print("N = 2^20: %.2f us/step" % step_us(1 << 20, stream, 200))
print("copy after the 2^20 handle: %.0f GB/s" % copy_gbs(stream))
for n in (1 << 22, 1 << 24):
    print("N = 2^%d after the 2^20 handle: %.1f us/step" % (n.bit_length() - 1, step_us(n, stream)))
