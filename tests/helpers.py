"""Shared test helpers: golden fixtures (reference outputs) and config handling."""
import ctypes as C
import json
import os

import numpy as np

from mbt_gym_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def golden_names():
    with open(os.path.join(GOLDEN_DIR, "index.json")) as f:
        return sorted(json.load(f).keys())


def golden_specs():
    with open(os.path.join(GOLDEN_DIR, "index.json")) as f:
        return json.load(f)


class Golden:
    """One fixture: the reference's outputs for a recorded action sequence (tools/make_golden.py)."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        raw = z["cfg"].tobytes()
        assert len(raw) == C.sizeof(_abi.mbt_config), "fixture was written with a different mbt_config layout"
        self.cfg = _abi.mbt_config.from_buffer_copy(raw)
        self.seed = int(z["seed"])
        self.n_episodes = int(z["n_episodes"])
        self.actions = z["actions"]        # (steps_total, N, A)
        self.reset_obs = z["reset_obs"]    # (n_episodes, N, D)
        self.obs = z["obs"]                # (steps_total, N, D)
        self.rew = z["rew"]                # (steps_total, N)
        self.done = z["done"].astype(bool)  # (steps_total,)
        self.final_state = z["final_state"]
        self.steps_per_episode = self.obs.shape[0] // self.n_episodes

    def config(self, precision=_abi.MBT_F64, **overrides):
        cfg = _abi.mbt_config.from_buffer_copy(bytes(self.cfg))
        cfg.precision = precision
        for k, v in overrides.items():
            setattr(cfg, k, v)
        return cfg

    # libm pow / exp (numpy) vs mbt_math differ by ulps in these two fixtures only
    @property
    def exact(self):
        return self.name not in ("rip_cubic", "exputil")


def copy_config(cfg, **overrides):
    out = _abi.mbt_config.from_buffer_copy(bytes(cfg))
    for k, v in overrides.items():
        setattr(out, k, v)
    return out


def assert_same(a, b, exact=True, rtol=1e-11, atol=1e-11, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if exact:
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            i = tuple(bad[0])
            raise AssertionError(f"{what}: {len(bad)} of {a.size} values differ; first at {i}: {a[i]!r} vs {b[i]!r}")
    else:
        np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, err_msg=what)
