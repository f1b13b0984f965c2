"""The kernels' per-trajectory core (mbt_gym_b200/csrc/mbt_step_core.cuh), compiled for the HOST by this test,
must reproduce the reference fixtures bit-for-bit in float64 and the oracle bit-for-bit in float32.

This is a build-container safety net (no GPU here); the same comparison through the real CUDA kernels and the
C ABI is tests/test_gpu_parity.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from mbt_gym_b200 import _abi
from oracle import oracle as O
from tests.helpers import Golden, ROOT, assert_same, golden_names

SRC = os.path.join(ROOT, "tests", "hostsim", "hostsim.cpp")
SO = os.path.join(ROOT, "tests", "hostsim", "_build", "libhostsim.so")


@pytest.fixture(scope="module")
def hostsim():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    cpuinfo = open("/proc/cpuinfo").read() if os.path.exists("/proc/cpuinfo") else ""
    march = ["-march=x86-64-v3"] if (" fma" in cpuinfo and " avx2" in cpuinfo) else []
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", *march, "-I", os.path.join(ROOT, "include"),
                    "-x", "c++", SRC, "-o", SO], check=True)
    lib = C.CDLL(SO)
    vp = C.c_void_p
    lib.hostsim_variant_of.argtypes = [C.POINTER(_abi.mbt_config)]
    lib.hostsim_run.argtypes = [C.POINTER(_abi.mbt_config), C.c_int, C.c_uint64, C.c_int64, C.c_double, C.c_double,
                                C.c_int, C.c_double, vp, vp, C.c_int, vp, vp, vp, vp]
    return lib


def run_hostsim(lib, g, precision, variant):
    """Oracle does the resets (state injection), the host-compiled kernel core does the steps."""
    cfg = g.config(precision)
    dt = np.float64 if precision == _abi.MBT_F64 else np.float32
    orc = O.OracleEnv(cfg)
    orc.seed(g.seed)
    N, A, D = orc.N, orc.A, orc.D
    spe = g.steps_per_episode
    obs_all, rew_all, done_all, oobs_all, orew_all = [], [], [], [], []
    for ep in range(g.n_episodes):
        orc.reset()
        state = orc.state.copy()
        q0 = np.ascontiguousarray(state[:, 1].copy())
        acts = np.ascontiguousarray(g.actions[ep * spe:(ep + 1) * spe], dtype=dt)
        obs = np.empty((spe, N, D), dt)
        rew = np.empty((spe, N), dt)
        dones = np.zeros(spe, np.uint8)
        per_traj = int(cfg.q0_mode == _abi.MBT_Q0_UNIFORM_INT)
        rc = lib.hostsim_run(C.byref(cfg), variant, g.seed, ep * spe, cfg.start_time, cfg.start_time, per_traj,
                             cfg.q0_const, state.ctypes.data, q0.ctypes.data, spe, acts.ctypes.data, obs.ctypes.data,
                             rew.ctypes.data, dones.ctypes.data)
        assert rc == 0
        obs_all.append(obs); rew_all.append(rew); done_all.append(dones.astype(bool))
        for k in range(spe):  # advance the oracle in lock-step (also gives the f32 comparison target)
            o, r, _ = orc.step(acts[k])
            oobs_all.append(o); orew_all.append(r)
    return (np.concatenate(obs_all), np.concatenate(rew_all), np.concatenate(done_all), np.stack(oobs_all),
            np.stack(orew_all))


@pytest.mark.parametrize("name", golden_names())
def test_kernel_core_f64_matches_reference_fixture(hostsim, name):
    g = Golden(name)
    for variant in sorted({0, hostsim.hostsim_variant_of(C.byref(g.cfg))}):
        obs, rew, done, _, _ = run_hostsim(hostsim, g, _abi.MBT_F64, variant)
        assert_same(obs, g.obs, exact=g.exact, what=f"{name} obs (variant {variant})")
        assert_same(rew, g.rew, exact=g.exact, what=f"{name} rew (variant {variant})")
        assert np.array_equal(done, g.done)


@pytest.mark.parametrize("name", golden_names())
def test_kernel_core_f32_matches_oracle_f32(hostsim, name):
    g = Golden(name)
    for variant in sorted({0, hostsim.hostsim_variant_of(C.byref(g.cfg))}):
        obs, rew, done, oobs, orew = run_hostsim(hostsim, g, _abi.MBT_F32, variant)
        assert_same(obs, oobs, exact=True, what=f"{name} f32 obs (variant {variant})")
        assert_same(rew, orew, exact=True, what=f"{name} f32 rew (variant {variant})")


@pytest.mark.parametrize("precision", [_abi.MBT_F64, _abi.MBT_F32])
def test_kernel_core_matches_oracle_on_random_configurations(hostsim, precision):
    """The same randomly drawn configurations the GPU test steps (tests/test_random_configs.py), through the host compile of
    the kernel core: every model kind, normalisation flag, start time and inventory mode against the oracle, bit-for-bit,
    in both precisions and through both the selected variant and the generic one."""
    from oracle import ref_shim as R  # make_actions only (no reference needed)
    from tests.helpers import build_facade_env
    from tests.test_random_configs import random_specs

    dt = np.float64 if precision == _abi.MBT_F64 else np.float32
    for spec in random_specs(60, 424242):
        env = build_facade_env(spec, precision="float64" if precision == _abi.MBT_F64 else "float32")
        cfg = env._build_config()
        acts = np.ascontiguousarray(R.make_actions(spec, env, spec["n_steps"], 11), dtype=dt)
        for variant in sorted({0, hostsim.hostsim_variant_of(C.byref(cfg))}):
            orc = O.OracleEnv(cfg)
            orc.seed(spec["seed"])
            orc.reset()
            state = orc.state.copy()
            q0 = np.ascontiguousarray(state[:, 1].copy())
            t0 = float(cfg.start_time)  # (the float32 state column holds a rounded copy; the clock itself is a double)
            N, D, n = orc.N, orc.D, spec["n_steps"]
            n_run = int(round((cfg.terminal_time - t0) / cfg.step_size))  # a late start shortens the episode
            n_run = max(1, min(n, n_run))
            obs = np.empty((n_run, N, D), dt)
            rew = np.empty((n_run, N), dt)
            dones = np.zeros(n_run, np.uint8)
            rc = hostsim.hostsim_run(C.byref(cfg), variant, spec["seed"], 0, t0, t0, int(cfg.q0_mode == _abi.MBT_Q0_UNIFORM_INT),
                                     cfg.q0_const, state.ctypes.data, q0.ctypes.data, n_run, acts.ctypes.data,
                                     obs.ctypes.data, rew.ctypes.data, dones.ctypes.data)
            assert rc == 0, spec
            for k in range(n_run):
                o, r, d = orc.step(acts[k])
                assert_same(obs[k], o, what=f"{spec} variant {variant} obs step {k}")
                assert_same(rew[k], r, what=f"{spec} variant {variant} rew step {k}")
                assert bool(dones[k]) == d
