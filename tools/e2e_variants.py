#!/usr/bin/env python
"""Host-buffer path variants at BASELINE size (1 GPU): chunk schedules of the H2D -> kernel -> D2H pipeline, zero-copy.
Each variant runs in a fresh process (the schedule is read once per process):  python tools/e2e_variants.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, time, json
sys.path.insert(0, %r)
import numpy as np
import bench
f = bench.make_env("as", "float64", 1 << 20, 0, 0)
a = f.pinned_actions(); a[:] = 0.7
f.reset()
for _ in range(8):
    o, r, d, i = f.step(a)
ts = []
for _ in range(60):
    t0 = time.perf_counter(); o, r, d, i = f.step(a); ts.append(time.perf_counter() - t0)
    if d[0]: f.reset()
# the C call alone (no Python facade work): native handle, pinned buffers
n = f._native
ob, rw = f._out_buffers()
tn = []
for _ in range(60):
    t0 = time.perf_counter(); done = n.step(a, ob, rw); tn.append(time.perf_counter() - t0)
    if done: n.reset()
print(json.dumps({"facade_ms_median": 1e3 * float(np.median(ts)), "facade_ms_mean": 1e3 * float(np.mean(ts)),
                  "native_ms_median": 1e3 * float(np.median(tn))}))
''' % ROOT

variants = [("geo (default)", {}), ("equal 4", {"MBT_PIPE_CHUNKS": "4"}), ("equal 8", {"MBT_PIPE_CHUNKS": "8"}),
            ("equal 2", {"MBT_PIPE_CHUNKS": "2"}), ("equal 1", {"MBT_PIPE_CHUNKS": "1"}), ("zerocopy", {"MBT_HOST_PATH": "zerocopy"})]
for name, env in variants:
    out = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, **env), capture_output=True, text=True)
    print(f"{name:16s}", out.stdout.strip() or out.stderr[-400:])
