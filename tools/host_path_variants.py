#!/usr/bin/env python
"""What a caller's choice of arrays costs on the host-buffer path (N = 2^20, float64): pinned actions (env.pinned_actions()),
an ordinary NumPy array reused every step, a FRESH NumPy array every step (what `agent.get_action(obs)` returns), and
copy_outputs=True.  python tools/host_path_variants.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def run(name, make_action, **kw):
    f = bench.make_env("as", "float64", 1 << 20, 0, 0)
    for k, v in kw.items():
        setattr(f, k, v)
    f.reset()
    for _ in range(6):
        o, r, d, i = f.step(make_action(f))
    ts = []
    for _ in range(40):
        a = make_action(f)
        t0 = time.perf_counter()
        o, r, d, i = f.step(a)
        ts.append(time.perf_counter() - t0)
        if d[0]:
            f.reset()
    print(f"{name:58s} {1e3 * np.median(ts):.3f} ms per step (median of 40)", flush=True)
    f.close()


pinned = {}
run("pinned actions (env.pinned_actions())", lambda f: pinned.setdefault(id(f), f.pinned_actions()))
reused = np.full((1 << 20, 2), 0.7)
run("one ordinary NumPy array, reused", lambda f: reused)
run("a fresh NumPy array every step", lambda f: np.full((1 << 20, 2), 0.7))
run("float32 actions (converted by the facade), fresh", lambda f: np.full((1 << 20, 2), 0.7, np.float32))
run("pinned actions, copy_outputs=True", lambda f: pinned.setdefault(id(f), f.pinned_actions()), copy_outputs=True)
