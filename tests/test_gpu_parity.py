"""GPU parity (-m gpu): the CUDA kernels, called through the C ABI, against
  (a) the committed REFERENCE fixtures (tests/golden, float64, bit-for-bit),
  (b) the CPU oracle on the same seeds (float32 bit-for-bit; float64 again),
  (c) size-independent properties at BASELINE.json's full size (2^20 trajectories).

Tolerances: float64 and float32 are compared BIT-EXACT (`np.array_equal`), except the two fixtures that reach
libm pow/exp in the reference (`rip_cubic`, `exputil`), compared at rtol=atol=1e-11.
"""
import numpy as np
import pytest

from mbt_gym_b200 import _abi, _lib
from oracle import oracle as O
from tests.helpers import Golden, assert_same, golden_names

pytestmark = pytest.mark.gpu


def run_native(g, precision, mem="host", n_shards=1):
    """Step the fixture's action sequence through libmbt_b200; returns (reset_obs, obs, rew, done, final_state)."""
    dt = np.float64 if precision == _abi.MBT_F64 else np.float32
    N_total = int(g.cfg.num_trajectories)
    bounds = np.linspace(0, N_total, n_shards + 1).astype(int)
    envs = []
    for s in range(n_shards):
        cfg = g.config(precision, num_trajectories=int(bounds[s + 1] - bounds[s]), traj_offset=int(bounds[s]))
        e = _lib.NativeEnv(cfg)
        e.seed(g.seed)
        envs.append(e)
    A, D = envs[0].A, envs[0].D
    spe = g.steps_per_episode
    reset_obs, obs_all, rew_all, done_all = [], [], [], []
    if mem == "device":
        import torch
    for ep in range(g.n_episodes):
        ro = np.empty((N_total, D), dt)
        for s, e in enumerate(envs):
            sl = slice(bounds[s], bounds[s + 1])
            buf = np.empty((e.N, D), dt)
            e.reset(buf)
            ro[sl] = buf
        reset_obs.append(ro)
        for k in range(ep * spe, (ep + 1) * spe):
            o = np.empty((N_total, D), dt)
            r = np.empty((N_total,), dt)
            dones = []
            for s, e in enumerate(envs):
                sl = slice(bounds[s], bounds[s + 1])
                a = np.ascontiguousarray(g.actions[k, sl], dtype=dt)
                if mem == "host":
                    ob = np.empty((e.N, D), dt)
                    rb = np.empty((e.N,), dt)
                    dones.append(e.step(a, ob, rb))
                else:
                    tdt = torch.float64 if precision == _abi.MBT_F64 else torch.float32
                    ta = torch.from_numpy(a).cuda()
                    to = torch.empty((e.N, D), dtype=tdt, device="cuda")
                    tr = torch.empty((e.N,), dtype=tdt, device="cuda")
                    torch.cuda.synchronize()
                    dones.append(e.step(ta, to, tr, mem=_abi.MBT_MEM_DEVICE))
                    e.sync()
                    ob, rb = to.cpu().numpy(), tr.cpu().numpy()
                o[sl], r[sl] = ob, rb
            assert len(set(dones)) == 1
            obs_all.append(o); rew_all.append(r); done_all.append(dones[0])
    final = np.concatenate([e.get_state() for e in envs])
    for e in envs:
        e.close()
    return np.stack(reset_obs), np.stack(obs_all), np.stack(rew_all), np.array(done_all), final


@pytest.mark.parametrize("name", golden_names())
def test_f64_matches_reference_fixture(name):
    g = Golden(name)
    reset_obs, obs, rew, done, final = run_native(g, _abi.MBT_F64)
    assert_same(reset_obs, g.reset_obs, exact=True, what=f"{name} reset obs")
    assert_same(obs, g.obs, exact=g.exact, what=f"{name} obs")
    assert_same(rew, g.rew, exact=g.exact, what=f"{name} rewards")
    assert np.array_equal(done, g.done)
    assert_same(final, g.final_state, exact=g.exact, what=f"{name} final state")


def oracle_run(g, precision):
    cfg = g.config(precision)
    orc = O.OracleEnv(cfg)
    orc.seed(g.seed)
    spe = g.steps_per_episode
    reset_obs, obs, rew, done = [], [], [], []
    for ep in range(g.n_episodes):
        reset_obs.append(orc.reset())
        for k in range(ep * spe, (ep + 1) * spe):
            o, r, d = orc.step(g.actions[k])
            obs.append(o); rew.append(r); done.append(d)
    return np.stack(reset_obs), np.stack(obs), np.stack(rew), np.array(done), orc.state


@pytest.mark.parametrize("name", golden_names())
def test_f32_matches_oracle_bitwise(name):
    g = Golden(name)
    got = run_native(g, _abi.MBT_F32)
    want = oracle_run(g, _abi.MBT_F32)
    for a, b, what in zip(got, want, ("reset obs", "obs", "rewards", "done", "final state")):
        assert_same(a, b, exact=True, what=f"{name} f32 {what}")


@pytest.mark.parametrize("name", ["as_pnl", "hawkes_normalised", "oe_ou_cjoe", "limit_and_market", "power_fill",
                                  "triangular_fill"])
@pytest.mark.parametrize("precision", [_abi.MBT_F64, _abi.MBT_F32])
def test_device_pointer_mode_equals_host_mode(name, precision):
    g = Golden(name)
    host = run_native(g, precision, mem="host")
    dev = run_native(g, precision, mem="device")
    for a, b, what in zip(host, dev, ("reset obs", "obs", "rewards", "done", "final state")):
        assert_same(a, b, exact=True, what=f"{name} {what}")


@pytest.mark.parametrize("name", ["as_pnl", "cjmm", "hawkes_pnl", "oe_ou_cjoe", "two_episodes"])
def test_results_do_not_depend_on_sharding(name):
    """Global trajectory ids in the Philox counter: 1, 2 or 3 handles give the same trajectories."""
    g = Golden(name)
    one = run_native(g, _abi.MBT_F64, n_shards=1)
    for shards in (2, 3):
        many = run_native(g, _abi.MBT_F64, n_shards=shards)
        for a, b, what in zip(one, many, ("reset obs", "obs", "rewards", "done", "final state")):
            assert_same(a, b, exact=True, what=f"{name} shards={shards} {what}")


def test_edge_sizes_and_errors():
    g = Golden("as_pnl")
    for n in (1, 2, 31, 33, 255, 257):
        cfg = g.config(_abi.MBT_F64, num_trajectories=n)
        e = _lib.NativeEnv(cfg)
        e.seed(3)
        orc = O.OracleEnv(cfg)
        orc.seed(3)
        ob = np.empty((n, 4)); rb = np.empty((n,))
        e.reset(ob)
        assert_same(ob, orc.reset(), what=f"n={n} reset")
        a = np.full((n, 2), 0.7)
        for _ in range(3):
            e.step(a, ob, rb)
            o, r, _d = orc.step(a)
            assert_same(ob, o, what=f"n={n} obs"); assert_same(rb, r, what=f"n={n} rew")
        e.close()
    # step before reset -> MBT_E_STATE ; bad config -> MBT_E_INVALID_ARG / UNSUPPORTED
    e = _lib.NativeEnv(g.config(_abi.MBT_F64))
    with pytest.raises(_lib.MbtError) as ei:
        e.step(np.zeros((e.N, 2)))
    assert ei.value.code == _abi.MBT_E_STATE
    e.close()
    with pytest.raises(_lib.MbtError) as ei:
        _lib.NativeEnv(g.config(_abi.MBT_F64, num_trajectories=0))
    assert ei.value.code == _abi.MBT_E_INVALID_ARG
    with pytest.raises(_lib.MbtError) as ei:
        _lib.NativeEnv(g.config(_abi.MBT_F64, dynamics=17))
    assert ei.value.code == _abi.MBT_E_UNSUPPORTED


def test_set_get_state_roundtrip_and_unaligned_buffers():
    g = Golden("hawkes_pnl")
    e = _lib.NativeEnv(g.config(_abi.MBT_F64))
    e.seed(9)
    e.reset()
    rng = np.random.default_rng(0)
    st = rng.normal(size=(e.N, e.D))
    st[:, 2] = 0.25
    e.set_state(st)
    assert_same(e.get_state(), st, what="state roundtrip")
    assert e.clock()["time"] == 0.25
    # deliberately misaligned host views (offset by one element) must still be handled
    big_a = np.zeros(e.N * e.A + 1); big_o = np.zeros(e.N * e.D + 1); big_r = np.zeros(e.N + 1)
    a = big_a[1:].reshape(e.N, e.A); a[:] = 0.5
    ob = big_o[1:].reshape(e.N, e.D); rb = big_r[1:]
    e.step(a, ob, rb)
    e2 = _lib.NativeEnv(g.config(_abi.MBT_F64))
    e2.seed(9); e2.reset(); e2.set_state(st)
    ob2 = np.empty((e.N, e.D)); rb2 = np.empty(e.N)
    e2.step(np.full((e.N, e.A), 0.5), ob2, rb2)
    assert_same(ob, ob2, what="misaligned obs"); assert_same(rb, rb2, what="misaligned rew")
    e.close(); e2.close()


# --------------------------------------------------------------------------- batch-reduced fill functions
@pytest.mark.parametrize("name", ["triangular_fill", "power_fill"])
@pytest.mark.parametrize("precision", [_abi.MBT_F64, _abi.MBT_F32])
def test_batch_fill_reduction_at_many_blocks(name, precision):
    """Triangular / Power fill functions (reference: np.max(depths, 0) over the TRAJECTORY axis): the reduction kernel
    in front of the step must find the deepest quote wherever it sits -- first row, last row, a row only the grid-stride
    loop reaches -- at a size that needs every block of its grid, propagate NaN like np.max, and match the oracle
    bit-for-bit.  N = 700 001 > 148 * 8 blocks * 256 threads, and not a multiple of anything."""
    g = Golden(name)
    dt = np.float64 if precision == _abi.MBT_F64 else np.float32
    N = 700_001
    cfg = g.config(precision, num_trajectories=N, normalise_action=0, normalise_obs=0)
    e = _lib.NativeEnv(cfg)
    orc = O.OracleEnv(cfg)
    e.seed(5); orc.seed(5)
    ob = np.empty((N, e.D), dt); rb = np.empty(N, dt)
    e.reset(ob)
    assert_same(ob, orc.reset(), what="reset")
    rng = np.random.default_rng(1)
    peaks = [0, N - 1, N // 2 + 17, 303_104 + 5, None, "nan"]
    for k, where in enumerate(peaks):
        a = rng.uniform(0.0, 0.4, size=(N, 2)).astype(dt)
        if where == "nan":
            a[N // 3, 1] = np.nan  # np.max -> NaN -> `unif < NaN` is False: nothing fills (on that side, or at all)
        elif where is not None:
            a[where, 0] = 0.55 + 0.05 * k
            a[(where * 7 + 3) % N, 1] = 0.75 - 0.05 * k
        e.step(a, ob, rb)
        o, r, _d = orc.step(a)
        assert_same(ob, o, what=f"{name} obs, peak at {where}")
        assert_same(rb, r, what=f"{name} rew, peak at {where}")
    assert len(np.unique(ob[:, 1])) > 3, "the fills must actually happen for the comparison to mean anything"
    e.close()


@pytest.mark.parametrize("precision", [_abi.MBT_F64, _abi.MBT_F32])
def test_batch_fill_fused_rollout_uniform_policies_only(precision):
    """With a batch-reduced fill function the fused rollout runs policies that are uniform over the batch (the deepest
    quote of the batch is then every trajectory's own) and refuses the state-dependent ones with MBT_E_UNSUPPORTED."""
    g = Golden("power_fill")
    dt = np.float64 if precision == _abi.MBT_F64 else np.float32
    cfg = g.config(precision, num_trajectories=777)
    e = _lib.NativeEnv(cfg)
    orc = O.OracleEnv(cfg)
    e.seed(8); orc.seed(8)
    e.reset(); orc.reset()
    pol = _abi.mbt_policy()
    pol.kind = _abi.MBT_POL_FIXED
    pol.fixed[0], pol.fixed[1] = -0.6, -0.4  # normalised action units (the fixture normalises actions)
    ret = np.empty(e.N, dt); qT = np.empty(e.N, dt)
    e.rollout(pol, ret, qT)
    R = np.zeros(e.N, dt)
    done = False
    a = np.tile(np.array([[-0.6, -0.4]], dt), (e.N, 1))
    while not done:
        _o, r, done = orc.step(a)
        R = R + r
    assert_same(ret, R, what="per-trajectory returns")
    assert_same(e.get_state(), orc.state, what="terminal state")
    assert len(np.unique(qT)) > 3
    e.reset()
    pol.kind = _abi.MBT_POL_AVELLANEDA_STOIKOV
    pol.as_gamma, pol.as_sigma_sq, pol.as_fill_comp, pol.as_terminal_time = 0.1, 4.0, 1.0, 1.0
    with pytest.raises(_lib.MbtError) as ei:
        e.rollout(pol, ret)
    assert ei.value.code == _abi.MBT_E_UNSUPPORTED
    e.close()


# --------------------------------------------------------------------------- fused rollout
def as_policy_numpy(state, gamma, sigma, kappa, T, dtype):
    """AvellanedaStoikovAgent.get_action (BaselineAgents.py:62-83), same evaluation order as the kernel."""
    q = state[:, 1].astype(dtype)
    t = state[:, 2].astype(dtype)
    g, s2 = dtype(gamma), dtype(sigma ** 2)
    fill = dtype(2 / gamma * np.log(1 + gamma / kappa))
    tau = dtype(T) - t
    adj = ((q * g) * s2) * tau
    spread = (g * s2) * tau + fill
    return np.stack([adj + spread / dtype(2), -adj + spread / dtype(2)], axis=1)


@pytest.mark.parametrize("precision", [_abi.MBT_F64, _abi.MBT_F32])
def test_fused_rollout_matches_stepwise_oracle(precision):
    g = Golden("as_pnl")
    dt = np.float64 if precision == _abi.MBT_F64 else np.float32
    cfg = g.config(precision, num_trajectories=1000)
    gamma = 0.1
    e = _lib.NativeEnv(cfg)
    e.seed(50)
    e.reset()
    pol = _abi.mbt_policy()
    pol.kind = _abi.MBT_POL_AVELLANEDA_STOIKOV
    pol.as_gamma, pol.as_sigma_sq = gamma, cfg.mid_vol ** 2
    pol.as_fill_comp = 2 / gamma * np.log(1 + gamma / cfg.fill_exponent)
    pol.as_terminal_time = cfg.terminal_time
    ret = np.empty(e.N, dt); qT = np.empty(e.N, dt)
    summ = e.rollout(pol, ret, qT)
    # oracle, stepped with the numpy policy, returns accumulated sequentially like the kernel does
    orc = O.OracleEnv(cfg)
    orc.seed(50)
    orc.reset()
    R = np.zeros(e.N, dt); act_sum = 0.0; r2 = 0.0
    done = False
    steps = 0
    while not done:
        a = as_policy_numpy(orc.state, gamma, cfg.mid_vol, cfg.fill_exponent, cfg.terminal_time, dt)
        act_sum += float(a.astype(np.float64).sum())
        _o, r, done = orc.step(a)
        R = R + r
        r2 += float((r.astype(np.float64) ** 2).sum())
        steps += 1
    assert summ.steps == steps == cfg.n_steps
    assert_same(ret, R, exact=True, what="per-trajectory returns")
    assert_same(qT, orc.state[:, 1], exact=True, what="terminal inventories")
    assert_same(e.get_state(), orc.state, exact=True, what="terminal state")
    R64 = R.astype(np.float64); q64 = orc.state[:, 1].astype(np.float64)
    np.testing.assert_allclose(summ.sum_return, R64.sum(), rtol=1e-12)
    np.testing.assert_allclose(summ.sum_return_sq, (R64 ** 2).sum(), rtol=1e-12)
    np.testing.assert_allclose(summ.sum_q, q64.sum(), rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(summ.sum_q_sq, (q64 ** 2).sum(), rtol=1e-12)
    np.testing.assert_allclose(summ.sum_action, act_sum, rtol=1e-9)
    np.testing.assert_allclose(summ.sum_reward_sq, r2, rtol=1e-9)
    e.close()


# --------------------------------------------------------------------------- full size (BASELINE configs[1])
def test_full_size_properties_2pow20():
    """N = 2^20: determinism, sharding invariance on a slice, and the AS-2008 statistics of the reference notebook
    (Test_1: mean PnL 64.872, std 6.69, mean q_T 0.201, std q_T 2.89 at N=1000, seed 50) within 4 standard errors."""
    g = Golden("as_pnl")
    N = 1 << 20
    cfg = g.config(_abi.MBT_F64, num_trajectories=N)
    gamma = 0.1
    pol = _abi.mbt_policy()
    pol.kind = _abi.MBT_POL_AVELLANEDA_STOIKOV
    pol.as_gamma, pol.as_sigma_sq = gamma, cfg.mid_vol ** 2
    pol.as_fill_comp = 2 / gamma * np.log(1 + gamma / cfg.fill_exponent)
    pol.as_terminal_time = cfg.terminal_time

    def run(cfg_):
        e = _lib.NativeEnv(cfg_)
        e.seed(50)
        e.reset()
        ret = np.empty(e.N); qT = np.empty(e.N)
        s = e.rollout(pol, ret, qT)
        e.close()
        return ret, qT, s

    ret, qT, s = run(cfg)
    ret2, qT2, _ = run(cfg)
    assert np.array_equal(ret, ret2) and np.array_equal(qT, qT2), "same seed must give identical trajectories"
    # a shard holding global ids [N/2, N/2 + 4096) reproduces that slice
    part = g.config(_abi.MBT_F64, num_trajectories=4096, traj_offset=N // 2)
    ret3, qT3, _ = run(part)
    assert np.array_equal(ret3, ret[N // 2:N // 2 + 4096]) and np.array_equal(qT3, qT[N // 2:N // 2 + 4096])
    # statistics vs the reference notebook goldens (SE of the notebook's N=1000 sample dominates)
    mean_pnl, std_pnl = ret.mean(), ret.std()
    assert abs(mean_pnl - 64.872139) < 4 * 6.692567 / np.sqrt(1000)
    assert abs(std_pnl - 6.692567) < 4 * 6.692567 / np.sqrt(2 * 1000)
    assert abs(qT.mean() - 0.201) < 4 * 2.893544 / np.sqrt(1000)
    assert abs(qT.std() - 2.893544) < 4 * 2.893544 / np.sqrt(2 * 1000)
    np.testing.assert_allclose(s.sum_return / N, mean_pnl, rtol=1e-10)
    assert np.all(qT == np.round(qT))


def _fixed_policy(values):
    pol = _abi.mbt_policy()
    pol.kind = _abi.MBT_POL_FIXED
    for j, v in enumerate(values):
        pol.fixed[j] = v
    return pol


def _episode(cfg, pol, seed=77):
    e = _lib.NativeEnv(cfg)
    e.seed(seed)
    e.reset()
    ret = np.empty(e.N); qT = np.empty(e.N)
    summ = e.rollout(pol, ret, qT)
    state = e.get_state()
    e.close()
    return ret, qT, state, summ


def test_full_size_properties_cjmm_hawkes_oe_2pow20():
    """BASELINE.json configs[2..4] at their full per-GPU size (2^20 trajectories), through size-independent properties:
      C3  the reference's own unit-test identity (rewards/tests/testRewardFunctions.py:68-135): over an episode the
          CjMmCriterion rewards sum to the RunningInventoryPenalty rewards -- random initial inventories, late start;
      C4  Hawkes: integral inventories, positive finite intensities, and a shard holding global ids [a, a+4096) reproduces
          that slice of the full batch bit-for-bit;
      C5  optimal execution (8 x 2^20 trajectories sharded over 8 GPUs): inventory and permanent impact under a constant
          trading speed are the same deterministic recurrences for every trajectory, and the LAST shard's slice
          (global ids 7 * 2^20 + ...) is reproduced by a small handle with the same offset."""
    N = 1 << 20
    # ---- C3
    g = Golden("cjmm")
    pol = _fixed_policy([0.7, 0.9])
    cfg_mm = g.config(_abi.MBT_F64, num_trajectories=N)
    cfg_rip = g.config(_abi.MBT_F64, num_trajectories=N, reward=_abi.MBT_REW_RUNNING_INVENTORY_PENALTY)
    r_mm, q_mm, st_mm, s_mm = _episode(cfg_mm, pol)
    r_rip, q_rip, st_rip, _ = _episode(cfg_rip, pol)
    assert np.array_equal(st_mm, st_rip), "the reward function must not influence the dynamics"
    assert len(np.unique(q_mm)) > 20 and s_mm.steps == 90  # start_time 0.1 of 100 steps
    np.testing.assert_allclose(r_mm, r_rip, rtol=0, atol=1e-9)  # telescoping identity, float64 accumulation error only
    # ---- C4
    g = Golden("hawkes_pnl")
    pol = _fixed_policy([0.7, 0.7])
    cfg = g.config(_abi.MBT_F64, num_trajectories=N)
    ret, qT, st, summ = _episode(cfg, pol)
    assert summ.steps == 200 and np.all(qT == np.round(qT))
    lam = st[:, 4:6]
    assert np.all(np.isfinite(lam)) and np.all(lam > 0) and 5.0 < lam.mean() < 60.0
    a0 = 3 * (N // 4) + 123
    part = g.config(_abi.MBT_F64, num_trajectories=4096, traj_offset=a0)
    ret_p, qT_p, st_p, _ = _episode(part, pol)
    assert np.array_equal(ret_p, ret[a0:a0 + 4096]) and np.array_equal(st_p, st[a0:a0 + 4096])
    # ---- C5 (the shard of the last of 8 GPUs: global ids 7 * 2^20 ...)
    g = Golden("oe_ou_cjoe")
    pol = _fixed_policy([-1.0])
    off = 7 * N
    cfg = g.config(_abi.MBT_F64, num_trajectories=N, traj_offset=off)
    ret, qT, st, summ = _episode(cfg, pol)
    q, imp = cfg.q0_const, 0.0
    for _ in range(summ.steps):  # ModelDynamics.py:262-267, price_impact_models.py:88-89 with nu = -1
        q = q + (-1.0 * cfg.mid_step)
        imp = imp + (cfg.imp_perm * -1.0) * cfg.imp_step
    assert np.all(st[:, 1] == q) and np.all(st[:, 4] == imp) and np.all(qT == q)
    assert len(np.unique(st[:, 3])) > N // 2 and np.all(np.isfinite(ret))
    part = g.config(_abi.MBT_F64, num_trajectories=2048, traj_offset=off + N - 2048)
    ret_p, _, st_p, _ = _episode(part, pol)
    assert np.array_equal(ret_p, ret[N - 2048:]) and np.array_equal(st_p, st[N - 2048:])


def test_inventory_distribution_and_reward_per_step_vs_numpy_port():
    """statistical parity of the Philox/CUDA path with the PCG64/NumPy path on the quantities
    BASELINE.json names -- terminal-inventory distribution (chi-square, p > 0.001), mean terminal PnL and mean reward per
    step (4 standard errors) -- for the Avellaneda-Stoikov agent.  The NumPy port reproduces the reference's notebook
    golden exactly (tests/test_reference_live.py), so it stands in for the reference on the GPU box."""
    from scipy import stats

    from oracle import numpy_port as P

    g = Golden("as_pnl")
    gamma, n_gpu, n_ref = 0.1, 1 << 18, 1 << 15
    cfg = g.config(_abi.MBT_F64, num_trajectories=n_gpu)
    pol = _abi.mbt_policy()
    pol.kind = _abi.MBT_POL_AVELLANEDA_STOIKOV
    pol.as_gamma, pol.as_sigma_sq = gamma, cfg.mid_vol ** 2
    pol.as_fill_comp = 2 / gamma * np.log(1 + gamma / cfg.fill_exponent)
    pol.as_terminal_time = cfg.terminal_time
    e = _lib.NativeEnv(cfg)
    e.seed(2024)
    e.reset()
    ret = np.empty(n_gpu); qT = np.empty(n_gpu)
    e.rollout(pol, ret, qT)
    e.close()
    ref = P.NumpyPortEnv("as", N=n_ref, seed=77)
    obs = ref.reset()
    R = np.zeros(n_ref)
    while True:
        obs, r, d, _ = ref.step(P.as_agent_action(obs, gamma, 2.0, 1.5, 1.0))
        R += r
        if d[0]:
            break
    q_ref = obs[:, 1]
    # mean terminal PnL and mean reward per step
    se = np.hypot(ret.std() / np.sqrt(n_gpu), R.std() / np.sqrt(n_ref))
    assert abs(ret.mean() - R.mean()) < 4 * se
    assert abs(ret.mean() / 200 - R.mean() / 200) < 4 * se / 200
    assert abs(ret.std() - R.std()) < 4 * R.std() / np.sqrt(2 * n_ref)
    # terminal inventory histogram: chi-square homogeneity test, bins with expected count >= 10
    lo, hi = int(min(qT.min(), q_ref.min())), int(max(qT.max(), q_ref.max()))
    bins = np.arange(lo - 0.5, hi + 1.5)
    h_gpu, _ = np.histogram(qT, bins)
    h_ref, _ = np.histogram(q_ref, bins)
    keep = (h_gpu + h_ref) * (n_ref / (n_gpu + n_ref)) >= 10
    table = np.array([np.append(h_gpu[keep], h_gpu[~keep].sum()), np.append(h_ref[keep], h_ref[~keep].sum())])
    table = table[:, table.sum(axis=0) > 0]
    chi2, p, dof, _ = stats.chi2_contingency(table)
    assert p > 0.001, (chi2, p, dof)


def test_handles_are_independent_across_threads():
    """include/mbt_b200.h: a handle is not thread-safe, but DIFFERENT handles may be driven from different threads
    (ctypes releases the GIL during the calls).  Four threads, four handles, results equal to the sequential run."""
    import threading

    names = ["as_pnl", "hawkes_pnl", "oe_ou_cjoe", "cjmm"]
    sequential = {n: run_native(Golden(n), _abi.MBT_F64) for n in names}
    results, errors = {}, []

    def work(n):
        try:
            results[n] = run_native(Golden(n), _abi.MBT_F64)
        except Exception as exc:  # noqa: BLE001
            errors.append((n, exc))

    threads = [threading.Thread(target=work, args=(n,)) for n in names]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for n in names:
        for a, b, what in zip(results[n], sequential[n], ("reset obs", "obs", "rewards", "done", "final state")):
            assert_same(a, b, what=f"{n} threaded {what}")


def test_create_destroy_does_not_leak_device_or_pinned_memory():
    import torch

    g = Golden("as_pnl")
    cfg = g.config(_abi.MBT_F64, num_trajectories=1 << 18)
    a = np.full((1 << 18, 2), 0.7)
    o = np.empty((1 << 18, 4)); r = np.empty(1 << 18)

    def cycle():
        e = _lib.NativeEnv(cfg)
        e.seed(1)
        e.reset(o)
        e.step(a, o, r)
        pol = _abi.mbt_policy()
        pol.kind = _abi.MBT_POL_FIXED
        pol.fixed[0] = pol.fixed[1] = 0.7
        e.rollout(pol, r)
        blob = e.checkpoint()
        e.restore(blob)
        e.close()

    cycle()
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(30):
        cycle()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 8 << 20, f"device memory shrank by {(free0 - free1) >> 20} MiB over 30 create/destroy cycles"
