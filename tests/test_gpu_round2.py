"""GPU tests (-m gpu) of what round 2 added or pinned: callable initial inventories / start times, `seed=0`, edits of
model attributes (in place mid-episode, handle rebuilds that continue the random streams), pooled output arrays, CUDA-graph
capture after eager use, clip-event counting, and the NCCL group of handles (one-rank group here; two ranks in
tests/test_gpu_group.py when two GPUs are visible)."""
import ctypes as C
import gc

import numpy as np
import pytest

from mbt_gym_b200 import _abi, _lib
from oracle import oracle as O
from tests.helpers import Golden, ROOT, assert_same, build_facade_env, golden_names, golden_specs

pytestmark = pytest.mark.gpu
SPECS = golden_specs()


def _oracle_for(env):
    """The oracle configured from the facade's own flattened config, seeded with the facade's key."""
    cfg = env._build_config()
    orc = O.OracleEnv(cfg)
    orc.seed(env._key)
    return orc


# ------------------------------------------------------------------ a19: callable initial_inventory / start_time
def test_callable_initial_inventory_and_start_time_follow_the_reference_semantics():
    """TradingEnvironment.py:257-281: a callable start time is called at every reset and quantised to the step grid; a
    callable initial inventory is called at every reset, rounded to an int for limit-order dynamics, and assigned to the
    whole inventory column."""
    spec = dict(SPECS["cjmm"], N=257, n_steps=20)
    env = build_facade_env(spec)
    starts = iter([0.1234, 0.52, 0.0])
    invs = iter([2.6, -3.2, 0.4])
    env.start_time = lambda: next(starts)
    env.initial_inventory = lambda: next(invs)
    orc = _oracle_for(env)
    acts = np.random.default_rng(3).uniform(0.1, 1.4, size=(40, 257, 2))
    k = 0
    for t_raw, q_raw in ((0.1234, 2.6), (0.52, -3.2), (0.0, 0.4)):
        t0 = np.round(t_raw / env.step_size) * env.step_size
        q0 = float(int(np.round(q_raw)))
        obs = env.reset()
        assert np.all(obs[:, 2] == t0) and np.all(obs[:, 1] == q0)
        assert env.reward_function.initial_inventory == q0 and env.reward_function.episode_length == env.terminal_time - t0
        args = _abi.mbt_reset_args(start_time=float(t0), q0_mode=_abi.MBT_Q0_CONST, q0_const=q0)
        assert_same(obs, orc.reset(args), what="reset obs")
        done = False
        while not done:
            o, r, d, _ = env.step(acts[k % 40])
            oo, orr, od = orc.step(acts[k % 40])
            assert_same(o, oo, what="obs"); assert_same(r, orr, what="CjMm rewards (use q0 and the episode length)")
            done = bool(d[0]); assert done == od
            k += 1
    env.close()


def test_callable_initial_inventory_returning_one_value_per_trajectory():
    """With speed dynamics (`round_initial_inventory = False`) the reference assigns an array returned by the callable to
    the inventory column as is (:137,275-279), and CjOeCriterion.reset captures it per trajectory (RewardFunctions.py:72)."""
    spec = dict(SPECS["oe_ou_cjoe"], N=300, n_steps=15)
    rng = np.random.default_rng(11)
    q0s = [rng.integers(20, 120, size=300).astype(float), rng.uniform(-50, 50, size=300)]
    for precision, dt in (("float64", np.float64), ("float32", np.float32)):
        env = build_facade_env(spec, precision=precision)
        it = iter(q0s)
        env.initial_inventory = lambda: next(it)
        orc = _oracle_for(env)
        for ep in range(2):
            obs = env.reset()
            assert_same(obs[:, 1], q0s[ep].astype(dt), what="inventory column = the callable's array")
            assert np.array_equal(env.reward_function.initial_inventory, q0s[ep])
            args = _abi.mbt_reset_args(start_time=0.0, q0_mode=_abi.MBT_Q0_PER_TRAJ, q0_values=q0s[ep].ctypes.data)
            assert_same(obs, orc.reset(args), what="reset obs")
            a = np.full((300, 1), -3.0, dt)
            done = False
            while not done:
                o, r, d, _ = env.step(a)
                oo, orr, od = orc.step(a)
                assert_same(o, oo, what="obs"); assert_same(r, orr, what="CjOe rewards with per-trajectory q0")
                done = bool(d[0])
        env.close()


# ------------------------------------------------------------------ a20: seed semantics
def test_seed_zero_is_unseeded_like_the_reference_and_seed_method_rekeys():
    """`if seed:` (TradingEnvironment.py:70): seed=0 / None leave the generators unseeded (fresh entropy per env); a
    non-zero seed, or calling env.seed(...) (also with 0, :345-348), makes runs reproducible."""
    spec = dict(SPECS["as_pnl"], N=512, n_steps=10)
    a = np.full((512, 2), 0.7)

    def episode(env):
        env.reset()
        return np.stack([env.step(a)[0].copy() for _ in range(10)])

    same = [episode(build_facade_env(dict(spec, seed=77))) for _ in range(2)]
    assert np.array_equal(same[0], same[1])
    for s in (0, None):
        runs = [episode(build_facade_env(dict(spec, seed=s))) for _ in range(2)]
        assert not np.array_equal(runs[0], runs[1]), f"seed={s!r} must not be reproducible"
    e1, e2 = build_facade_env(dict(spec, seed=None)), build_facade_env(dict(spec, seed=None))
    e1.seed(0); e2.seed(0)
    assert np.array_equal(episode(e1), episode(e2)), "env.seed(0) seeds (only the constructor ignores 0)"
    e1.seed(5)
    x = episode(e1)
    e1.seed(5)
    assert np.array_equal(x, episode(e1)), "seed() restarts the streams"


# ------------------------------------------------------------------ edits of attributes (handle rebuild)
def test_attribute_edit_mid_episode_takes_effect_at_the_next_step_without_losing_state():
    """The reference reads Python attributes at every step: editing max_inventory or a reward parameter mid-episode
    changes the NEXT step and nothing else.  Here the live handle is reconfigured in place (mbt_reconfigure)."""
    spec = dict(SPECS["rip"], N=400, n_steps=30, max_inventory=50)
    env = build_facade_env(spec)
    orc = _oracle_for(env)
    acts = np.random.default_rng(5).uniform(0.05, 0.4, size=(30, 400, 2))
    assert_same(env.reset(), orc.reset(), what="reset")
    for k in range(30):
        if k == 10:
            env.max_inventory = 2
            env.reward_function.per_step_inventory_aversion = 0.5
            cfg = env._build_config()
            state, clk = orc.state, orc.clock()
            orc = O.OracleEnv(cfg)  # the oracle has no in-place edit: a new one, put where the old one was
            orc.seed(env._key)
            orc.reset()
            orc.set_state(state)
            # (set_state moves the oracle's clock; its step counter restarts, so inject the draws of step k directly)
        o, r, d, _ = env.step(acts[k])
        if k >= 10:
            u, z = O.draws(_abi.MBT_F64, env._key, 0, 400, k)
            oo, orr, od = orc.step_draws(acts[k], u, z)
        else:
            oo, orr, od = orc.step(acts[k])
        assert_same(o, oo, what=f"obs step {k}"); assert_same(r, orr, what=f"rew step {k}")
    assert np.abs(env.state[:, 1]).max() <= 2
    env.close()


def test_shape_changing_edit_is_refused_mid_episode_and_rebuild_between_episodes_continues_the_streams():
    spec = dict(SPECS["as_pnl"], N=300, n_steps=8)
    a = np.full((300, 2), 0.7)
    ref = build_facade_env(spec)          # never edited: episodes 0, 1, 2 of the seed's streams
    want = []
    for _ in range(3):
        ref.reset()
        want.append(np.stack([ref.step(a)[0].copy() for _ in range(8)]))
    env = build_facade_env(spec)
    env.reset()
    got0 = [env.step(a)[0].copy() for _ in range(3)]
    env.num_trajectories = 200
    with pytest.raises(RuntimeError, match="middle of an episode"):
        env.step(a[:200])
    with pytest.raises(RuntimeError, match="middle of an episode"):
        env.state
    env.num_trajectories = 300            # back: nothing was lost
    got0 += [env.step(a)[0].copy() for _ in range(5)]
    assert np.array_equal(np.stack(got0), want[0])
    # between episodes: a new handle (another num_trajectories = another state shape) inherits key and counters
    env.num_trajectories = 150
    obs = env.reset()
    assert obs.shape == (150, 4)
    got1 = np.stack([env.step(a[:150])[0].copy() for _ in range(8)])
    assert np.array_equal(got1, want[1][:, :150]), "the rebuilt handle continues the draw counters (episode 1, steps 8..15)"
    clk = env._native.clock()
    assert clk["n_step"] == 16 and clk["n_episode"] == 2
    env.close(); ref.close()


def test_checkpoint_restore_keeps_the_key_across_a_later_rebuild():
    spec = dict(SPECS["as_pnl"], N=128, n_steps=6)
    a = np.full((128, 2), 0.7)
    env = build_facade_env(spec)
    env.reset()
    env.step(a)
    blob = env.save_checkpoint()
    want = np.stack([env.step(a)[0].copy() for _ in range(5)])
    other = build_facade_env(dict(spec, seed=4242))
    other.load_checkpoint(blob)
    assert other._key == env._key
    other.max_cash = other.max_cash * 2   # an in-place edit after the restore must not re-seed with a stale key
    got = np.stack([other.step(a)[0].copy() for _ in range(5)])
    assert np.array_equal(got, want)
    env.close(); other.close()


# ------------------------------------------------------------------ pooled outputs (ring of 4)
def test_outputs_are_never_overwritten_while_referenced_and_outlive_close():
    g = Golden("as_pnl")
    env = build_facade_env(SPECS["as_pnl"])
    first = env.reset()
    kept, copies, views = [first], [first.copy()], []
    for k in range(12):  # far more than any ring
        o, r, d, _ = env.step(g.actions[k])
        assert not d.flags.writeable or d.base is None  # a fresh dones array
        kept += [o, r]; copies += [o.copy(), r.copy()]
        if k < 3:
            views.append((o[:, 3], o[:, 3].copy()))
        del o, r
    for arr, cp in zip(kept, copies):
        assert np.array_equal(arr, cp)
    n_blocks = len(env._pool._blocks)
    del kept, arr
    gc.collect()
    for _ in range(6):  # everything released except three column views: their blocks stay busy, the others are reused
        env.step(g.actions[0])
    assert len(env._pool._blocks) <= n_blocks
    env.close()
    gc.collect()
    for v, cp in views:
        assert np.array_equal(v, cp), "a view of a returned array stays valid after close()"
    # steady state: a caller that drops each step's arrays keeps reusing two blocks per shape
    env = build_facade_env(SPECS["as_pnl"])
    env.reset()
    for k in range(20):
        o, r, d, _ = env.step(g.actions[k])
    assert len(env._pool._blocks) <= 5
    env.close()


# ------------------------------------------------------------------ clip events: step loop == fused rollout
def test_clip_events_are_counted_alike_by_step_and_rollout():
    spec = dict(SPECS["as_pnl_tight"], N=2048, n_steps=40, max_cash=50.0)  # any open position breaks the cash bound
    cfg = build_facade_env(spec)._build_config()
    pol = _abi.mbt_policy()
    pol.kind = _abi.MBT_POL_FIXED
    pol.fixed[0], pol.fixed[1] = 0.05, 0.05   # both sides fill almost always
    a = np.tile(np.array([[0.05, 0.05]]), (2048, 1))
    env = _lib.NativeEnv(cfg); env.seed(9); env.reset()
    done = False
    while not done:
        done = env.step(a)
    stepped = env.clip_count()
    env.close()
    env = _lib.NativeEnv(cfg); env.seed(9); env.reset()
    summ = env.rollout(pol)
    assert summ.clipped == stepped == env.clip_count() and stepped > 2048
    env.close()


# ------------------------------------------------------------------ CUDA graphs after eager use (fold_counters)
def test_cuda_graph_capture_after_warmup_and_eager_calls_between_replays():
    import torch

    spec = dict(SPECS["cjmm"], N=3000, n_steps=6, seed=123, start_time=0.0)
    a_t = torch.full((3000, 2), 0.6, dtype=torch.float64, device="cuda")

    def eager_episode(env):
        env.reset_device()
        ret = torch.zeros(3000, dtype=torch.float64, device="cuda")
        for _ in range(6):
            _o, rew, _d, _ = env.step(a_t)
            ret = ret + rew
        torch.cuda.synchronize()
        return ret.cpu().numpy()

    ref = build_facade_env(spec)
    want = [eager_episode(ref) for _ in range(6)]
    ref.close()

    env = build_facade_env(spec)
    s = torch.cuda.Stream()
    ret_t = torch.zeros(3000, dtype=torch.float64, device="cuda")
    with torch.cuda.stream(s):
        got = [eager_episode(env)]            # a warm-up episode BEFORE the capture: consumes episode 0
    assert_same(got[0], want[0])
    torch.cuda.synchronize()
    # capturing now without prepare_capture() must fail loudly, not skip draw indices silently
    g_bad = torch.cuda.CUDAGraph()
    with pytest.raises(Exception):
        with torch.cuda.graph(g_bad, stream=s):
            env.reset_device()
    torch.cuda.synchronize()
    env.prepare_capture()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        env.reset_device()
        ret_t.zero_()
        for _ in range(6):
            _o, rew, _d, _ = env.step(a_t)
            ret_t += rew
        env.fold_counters()
    with torch.cuda.stream(s):
        got.append(eager_episode(env))        # eager episode between the capture and the first replay: episode 1
    g.replay(); torch.cuda.synchronize(); got.append(ret_t.cpu().numpy())   # episode 2
    g.replay(); torch.cuda.synchronize(); got.append(ret_t.cpu().numpy())   # episode 3
    with torch.cuda.stream(s):
        got.append(eager_episode(env))        # episode 4
    g.replay(); torch.cuda.synchronize(); got.append(ret_t.cpu().numpy())   # episode 5
    for i in range(6):
        assert_same(got[i], want[i], what=f"episode {i} (eager / replay interleaved)")
    clk = env._native.clock()
    assert clk["n_step"] == 36 and clk["n_episode"] == 6
    env.close()


# ------------------------------------------------------------------ group of handles (one rank)
def test_group_of_one_equals_plain_rollout_and_gathers_own_returns():
    import torch

    spec = dict(SPECS["oe_ou_cjoe"], N=5000, n_steps=20)
    cfg = build_facade_env(spec)._build_config()
    pol = _abi.mbt_policy()
    pol.kind = _abi.MBT_POL_FIXED
    pol.fixed[0] = -1.0
    plain = _lib.NativeEnv(cfg); plain.seed(3); plain.reset()
    ret = np.empty(5000)
    want = plain.rollout(pol, ret)
    plain.close()
    env = _lib.NativeEnv(cfg); env.seed(3)
    env.group_create(_lib.NativeEnv.group_unique_id(), 0, 1)
    assert env.group_info() == dict(rank=0, world=1, total_trajectories=5000)
    env.reset(mem=_abi.MBT_MEM_DEVICE)
    loc = torch.empty(5000, dtype=torch.float64, device="cuda")
    allr = torch.empty(5000, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    got = env.group_rollout(pol, loc, allr)
    env.group_wait()
    for f, _t in _abi.mbt_summary._fields_:
        assert getattr(got, f) == getattr(want, f), f
    assert np.array_equal(loc.cpu().numpy(), ret) and np.array_equal(allr.cpu().numpy(), ret)
    again = env.group_summary(want)
    assert again.sum_return == want.sum_return and again.count == 5000
    env.group_destroy()
    env.close()


# ------------------------------------------------------------------ draw contract: the tails of the midprice normal
def test_midprice_increments_have_normal_tails_at_full_size():
    """The draw contract (include/mbt_philox.h, DESIGN.md "Draw contract") turns 32 random bits into a normal by inversion
    in float32 arithmetic (2^31 probability levels, |z| <= 6.3) where the reference uses numpy's 53-bit ziggurat.  A
    tail-sensitive check at BASELINE size: 2^20 trajectories x 200 steps = 2.1e8 midprice increments, standardised --
    mean, variance, EXCESS KURTOSIS, the mass beyond 4 sigma and the largest |z| against the normal law.
    Tolerances: 5 standard errors of each statistic (SE(kurtosis) = sqrt(24/n) = 3.4e-4)."""
    import torch

    from scipy import stats

    spec = dict(SPECS["as_pnl"], N=1 << 20, n_steps=200)
    env = build_facade_env(spec)
    sigma_sqdt = 2.0 * np.sqrt(1.0 / 200)
    a = torch.full((1 << 20, 2), 0.7, dtype=torch.float64, device="cuda")
    prev = env.reset_device()[:, 3].clone()
    s1 = s2 = s4 = 0.0
    beyond4 = 0
    zmax = 0.0
    for _ in range(200):
        obs, _r, _d, _ = env.step(a)
        z = (obs[:, 3] - prev) / sigma_sqdt
        prev = obs[:, 3].clone()
        z2 = z * z
        s1 += float(z.sum()); s2 += float(z2.sum()); s4 += float((z2 * z2).sum())
        beyond4 += int((z.abs() > 4.0).sum()); zmax = max(zmax, float(z.abs().max()))
    env.close()
    n = 200 * (1 << 20)
    mean, var = s1 / n, s2 / n
    kurt = (s4 / n) / var ** 2 - 3.0  # (mean is ~1e-5: central and raw moments agree far below the tolerance)
    assert abs(mean) < 5 / np.sqrt(n)
    assert abs(var - 1.0) < 5 * np.sqrt(2.0 / n)
    assert abs(kurt) < 5 * np.sqrt(24.0 / n), kurt
    p4 = 2 * stats.norm.sf(4.0)
    assert abs(beyond4 - n * p4) < 5 * np.sqrt(n * p4), (beyond4, n * p4)
    assert 5.3 < zmax <= 6.31, zmax  # P(max |z| < 5.3) = exp(-24); the contract's cut-off is 6.3


# ------------------------------------------------------------------ run-time specialised kernels (mbt_jit.h)
@pytest.mark.parametrize("name", golden_names())
def test_every_fixture_configuration_runs_a_kernel_without_model_switches(name):
    """Either one of the fully specialised ahead-of-time variants, or a kernel compiled at run time for exactly this
    configuration: no local memory, at most 50 registers (the generic kernel has 64).  The parity tests of this suite run
    through these kernels (tests/conftest.py sets MBT_JIT=require)."""
    g = Golden(name)
    for prec in (_abi.MBT_F64, _abi.MBT_F32):
        env = _lib.NativeEnv(g.config(prec))
        info = env.kernel_info()
        env.close()
        assert info["jit_mode"] == 2 and info["message"] == ""
        if info["aot_variant"] in (0, 4, 6, 9):  # generic / runtime-flag variants of the ahead-of-time table
            assert info["step_is_jit"], info
            assert info["step_local_bytes"] == 0 and 0 < info["step_registers"] <= 50, info
        else:
            assert not info["step_is_jit"]


def test_generic_ahead_of_time_kernels_still_match_the_reference_when_the_specialiser_is_off():
    """MBT_JIT=0 (no NVRTC on the machine): every configuration falls back to the ahead-of-time table -- the generic kernel
    for most fixtures -- and must still reproduce the reference fixtures bit-for-bit."""
    import os
    import subprocess
    import sys

    out = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-q", "-x", "-k",
                          "f64_matches_reference_fixture or f32_matches_oracle_bitwise or fused_rollout_matches"],
                         cwd=ROOT, env=dict(os.environ, MBT_JIT="0"), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:]
    assert " passed" in out.stdout


def test_reconfigure_switches_the_specialised_kernel():
    """An in-place edit that changes a model KIND (here the reward function) re-selects the kernel."""
    g = Golden("gbm_nonlinear")
    env = _lib.NativeEnv(g.config(_abi.MBT_F64))
    h0 = env.kernel_info()["jit_hash"]
    cfg = g.config(_abi.MBT_F64, reward=_abi.MBT_REW_RUNNING_INVENTORY_PENALTY, rew_phi=0.01, rew_alpha=0.1)
    env.reconfigure(cfg)
    info = env.kernel_info()
    assert info["step_is_jit"] and info["jit_hash"] != h0
    orc = O.OracleEnv(cfg)
    env.seed(5); orc.seed(5)
    obs = np.empty((env.N, env.D)); rew = np.empty(env.N)
    env.reset(obs)
    assert_same(obs, orc.reset())
    a = np.full((env.N, 2), 0.4)
    for _ in range(5):
        env.step(a, obs, rew)
        o, r, _d = orc.step(a)
        assert_same(obs, o); assert_same(rew, r)
    env.close()


# ------------------------------------------------------------------ inventory distribution on the device
def test_inventory_histogram_equals_numpy_bincount():
    """north_star's "inventory distribution" / the reference's results histogram (helpers/plotting.py:94-110): binned on the
    device from the inventory column, equal to numpy's count of the terminal inventories -- also with bins that cut the range
    (below / above counters) and for float32 state."""
    from mbt_gym_b200.agents.BaselineAgents import FixedSpreadAgent

    for precision in ("float64", "float32"):
        env = build_facade_env(dict(SPECS["as_pnl"], N=200_000, n_steps=60, max_inventory=60), precision=precision)
        env.reset()
        summ, ret, q = env.rollout_summary(FixedSpreadAgent(env, half_spread=0.6).to_policy(env), return_trajectory_stats=True)
        inv, counts, (below, above) = env.inventory_histogram()
        assert below == 0 and above == 0 and counts.sum() == 200_000
        want = np.array([(q == v).sum() for v in inv])
        assert np.array_equal(counts, want)
        assert abs((inv * counts).sum() - summ.sum_q) < 1e-6
        inv, counts, (below, above) = env.inventory_histogram(lo=-2, hi=3)
        assert below == (q < -2).sum() and above == (q > 3).sum() and np.array_equal(counts, [(q == v).sum() for v in range(-2, 4)])
        with pytest.raises(_lib.MbtError):
            env._native.inventory_histogram(0, 5000)
        env.close()


# ------------------------------------------------------------------ pageable caller arrays through the chunked pipeline
def test_pageable_arrays_through_the_chunked_pipeline_equal_the_pinned_path():
    """Ordinary NumPy action arrays and `copy_outputs=True` outputs are staged / unstaged chunk by chunk inside the
    H2D -> kernel -> D2H pipeline (N >= 2^17 uses several chunks): same results as pinned buffers, for a row count that
    is not a multiple of the chunk granularity."""
    n = (1 << 18) + 12345
    spec = dict(SPECS["hawkes_pnl"], N=n, n_steps=6)
    rng = np.random.default_rng(0)
    acts = rng.uniform(0.1, 1.5, size=(6, n, 2))
    pinned_env = build_facade_env(spec)
    pageable_env = build_facade_env(spec, copy_outputs=True)
    a_pin = pinned_env.pinned_actions()
    o1, o2 = pinned_env.reset(), pageable_env.reset()
    assert o2.flags.owndata and np.array_equal(o1, o2)
    for k in range(6):
        a_pin[:] = acts[k]
        o1, r1, d1, _ = pinned_env.step(a_pin)
        o2, r2, d2, _ = pageable_env.step(acts[k].copy())
        assert o2.flags.owndata and r2.flags.owndata
        assert np.array_equal(o1, o2) and np.array_equal(r1, r2) and d1[0] == d2[0]
    pinned_env.close(); pageable_env.close()
