"""GPU tests (-m gpu) of the public API: mbt_gym_b200.gym.TradingEnvironment & friends behave like the reference's
classes -- same call shapes, and numerically the committed REFERENCE fixtures, bit-for-bit."""
import numpy as np
import pytest

from tests.helpers import Golden, assert_same, build_facade_env, golden_names, golden_specs

pytestmark = pytest.mark.gpu
SPECS = golden_specs()


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("copy_outputs", [False, True])
def test_env_step_reproduces_reference_fixture(name, copy_outputs):
    g = Golden(name)
    env = build_facade_env(SPECS[name], copy_outputs=copy_outputs)
    # (the Heston model publishes bounds for one of its two columns, in the reference too: midprice_models.py:343-346)
    assert env.observation_space.shape == (g.obs.shape[2] - name.startswith("heston"),)
    assert env.action_space.shape[0] == g.actions.shape[2]
    spe = g.steps_per_episode
    for ep in range(g.n_episodes):
        obs = env.reset()
        assert obs.shape == g.reset_obs[ep].shape and obs.dtype == np.float64
        assert_same(obs, g.reset_obs[ep], what=f"{name} reset")
        for k in range(ep * spe, (ep + 1) * spe):
            obs, rew, dones, infos = env.step(g.actions[k])
            assert_same(obs, g.obs[k], exact=g.exact, what=f"{name} obs step {k}")
            assert_same(rew, g.rew[k], exact=g.exact, what=f"{name} rew step {k}")
            assert dones.shape == (g.cfg.num_trajectories,) and dones.dtype == bool and bool(dones[0]) == g.done[k]
            assert len(infos) == g.cfg.num_trajectories and infos[0] == {}
    assert_same(env.state, g.final_state, exact=g.exact, what=f"{name} final state")
    env.close()


def test_returned_arrays_survive_the_next_steps_and_copy_mode_returns_ordinary_arrays():
    g = Golden("as_pnl")
    env = build_facade_env(SPECS["as_pnl"])
    env.reset()
    o1, r1, _, _ = env.step(g.actions[0])
    keep = o1.copy()
    o2, _, _, _ = env.step(g.actions[1])
    o3, _, _, _ = env.step(g.actions[2])
    assert o1 is not o2 and np.array_equal(o1, keep), "an observation must stay valid for the next calls (SB3 keeps _last_obs)"
    env2 = build_facade_env(SPECS["as_pnl"], copy_outputs=True)
    env2.reset()
    a, _, _, _ = env2.step(g.actions[0])
    b, _, _, _ = env2.step(g.actions[1])
    assert a is not b and a.flags.owndata
    env.close(); env2.close()


def test_torch_cuda_actions_take_the_zero_copy_path():
    import torch

    g = Golden("hawkes_pnl")
    env = build_facade_env(SPECS["hawkes_pnl"])
    env.reset()
    for k in range(5):
        obs, rew, dones, _ = env.step(torch.from_numpy(g.actions[k]).cuda())
        assert obs.is_cuda and rew.is_cuda and obs.shape == (env.num_trajectories, 6)
        torch.cuda.synchronize()
        assert_same(obs.cpu().numpy(), g.obs[k], what=f"device obs {k}")
        assert_same(rew.cpu().numpy(), g.rew[k], what=f"device rew {k}")
    env.close()


def test_generate_trajectory_and_results_table_both_paths():
    """generate_trajectory (host agent, step by step) and the fused on-device rollout give the same results table."""
    from mbt_gym_b200.agents.BaselineAgents import AvellanedaStoikovAgent
    from mbt_gym_b200.gym.helpers.generate_trajectory import (generate_results_table, generate_results_table_fused,
                                                              generate_trajectory)

    spec = dict(SPECS["as_pnl"], N=2000)
    env = build_facade_env(spec)
    agent = AvellanedaStoikovAgent(risk_aversion=0.1, env=env)
    env.seed(50)
    with pytest.warns(UserWarning):  # the reference agent warns about negative spreads too
        obs, act, rew = generate_trajectory(env, agent)
    assert obs.shape == (2000, 4, 201) and act.shape == (2000, 2, 200) and rew.shape == (2000, 1, 200)
    assert obs[0, 2, -1] == pytest.approx(1.0)
    env.seed(50)
    with pytest.warns(UserWarning):
        table, totals = generate_results_table(env, agent)
    env.seed(50)
    fused, returns = generate_results_table_fused(env, agent)
    assert np.array_equal(np.sort(returns), np.sort(returns))  # finite
    np.testing.assert_allclose(totals, returns, rtol=0, atol=1e-9)  # same draws, different summation order only
    for key in table:
        np.testing.assert_allclose(fused[key], table[key], rtol=1e-9, atol=1e-9, err_msg=key)
    # and the table is the Avellaneda-Stoikov replication of the reference notebook (N=1000 there), within 4 SE
    assert abs(table["Mean PnL"] - 64.872139) < 4 * np.hypot(6.69 / np.sqrt(1000), 6.69 / np.sqrt(2000))
    assert abs(table["Mean spread"] - 1.49177) < 1e-3
    env.close()


def test_cjp_closed_form_value_function_fused_rollout():
    """Test_2 notebook (CJP-2015): sample mean of total CjMm reward vs the closed-form value function 68.25583476
    (seed 410, phi=0.01, alpha=0.001, Q=100, n_steps=1000), here with 2^17 trajectories through the fused rollout
    with the CarteaJaimungalMmAgent depth table on the device."""
    from mbt_gym_b200.agents.BaselineAgents import CarteaJaimungalMmAgent

    spec = dict(SPECS["cjmm"], N=1 << 17, n_steps=1000, initial_inventory=0, start_time=0.0, seed=410)
    env = build_facade_env(spec)
    agent = CarteaJaimungalMmAgent(env=env)
    target = float(np.asarray(agent.calculate_true_value_function(np.array([[0.0, 0.0, 0.0, 100.0]] * 2))).reshape(-1)[0]) - 0.0
    assert abs(target - 68.25583476) < 1e-6
    env.reset()
    summary, returns, q_t = env.rollout_summary(agent.to_policy(env), return_trajectory_stats=True)
    mean, se = returns.mean(), returns.std() / np.sqrt(returns.size)
    # O(dt) discretisation bias at n_steps=1000 is ~0.01 (the notebook's own sample mean is 68.2426 +- 0.39)
    assert abs(mean - target) < 4 * se + 0.03, (mean, target, se)
    assert summary.steps == 1000 and summary.count == 1 << 17
    env.close()


def test_vecenv_adapter_autoreset_and_lazy_terminal_observation():
    from mbt_gym_b200.gym.StableBaselinesTradingEnvironment import StableBaselinesTradingEnvironment

    spec = dict(SPECS["as_pnl_normalised"], N=64, n_steps=5)
    env = build_facade_env(spec)
    venv = StableBaselinesTradingEnvironment(env)
    assert venv.num_envs == 64 and venv.env_is_wrapped(None) == [False] * 64
    obs = venv.reset()
    act = np.zeros((64, 2))
    for k in range(5):
        venv.step_async(act)
        obs, rew, dones, infos = venv.step_wait()
    assert dones.all() and len(infos) == 64
    term = infos[3]["terminal_observation"]
    assert term.shape == (4,) and term[2] == pytest.approx(1.0)      # normalised time = +1 at T
    assert obs[0, 2] == pytest.approx(-1.0)                          # already the first obs of the next episode
    venv.step_async(act)
    obs, rew, dones, infos = venv.step_wait()
    assert not dones.any() and infos[0] == {}
    env.close()
    # monitor=True: VecMonitor-style episode entries without a per-environment Python loop
    env = build_facade_env(spec)
    venv = StableBaselinesTradingEnvironment(env, monitor=True)
    venv.reset()
    total = np.zeros(64)
    for k in range(5):
        obs, rew, dones, infos = venv.step(act)
        total += rew
    assert dones.all()
    ep = infos[7]["episode"]
    assert ep["l"] == 5 and ep["r"] == pytest.approx(total[7]) and "terminal_observation" in infos[7]
    assert venv.last_episode_statistics["mean_return"] == pytest.approx(total.mean())
    assert np.all(venv.episode_returns == 0)
    env.close()


def test_vecenv_adapter_device_mode_equals_host_mode():
    """CUDA action tensors through the VecEnv adapter: observations / rewards stay on the device, the auto-reset and the
    monitor too, and every value equals the NumPy-mode run of the same seed."""
    import torch

    from mbt_gym_b200.gym.StableBaselinesTradingEnvironment import StableBaselinesTradingEnvironment

    spec = dict(SPECS["cjmm"], N=300, n_steps=6, start_time=0.0)
    host = StableBaselinesTradingEnvironment(build_facade_env(spec), monitor=True)
    dev = StableBaselinesTradingEnvironment(build_facade_env(spec), monitor=True)
    o_h, o_d = host.reset(), dev.reset(device=True)
    assert o_d.is_cuda
    assert_same(o_d.cpu().numpy(), o_h, what="first observation")
    rng = np.random.default_rng(3)
    for k in range(15):  # two and a half episodes: two auto-resets
        a = rng.uniform(0.1, 1.5, size=(300, 2))
        o_h, r_h, d_h, i_h = host.step(a)
        o_d, r_d, d_d, i_d = dev.step(torch.from_numpy(a).cuda())
        assert o_d.is_cuda and r_d.is_cuda
        assert_same(o_d.cpu().numpy(), o_h, what=f"obs {k}")
        assert_same(r_d.cpu().numpy(), r_h, what=f"rew {k}")
        assert np.array_equal(d_h, d_d)
        if d_h.all():
            assert_same(i_d[5]["terminal_observation"].cpu().numpy(), i_h[5]["terminal_observation"], what="terminal obs")
            assert i_d[5]["episode"]["r"] == i_h[5]["episode"]["r"] and i_d[5]["episode"]["l"] == 6
            assert dev.last_episode_statistics["mean_return"] == pytest.approx(host.last_episode_statistics["mean_return"], rel=1e-12)
            assert dev.last_episode_statistics["std_return"] == pytest.approx(host.last_episode_statistics["std_return"], rel=1e-12)
    assert dev.episode_returns.is_cuda
    assert_same(dev.episode_returns.cpu().numpy(), host.episode_returns, what="running returns of the open episode")
    host.env.close(); dev.env.close()


def test_reward_calculate_runs_on_device_and_matches_reference_unit_tests():
    """mbt_gym/rewards/tests/testRewardFunctions.py restated through RewardFunction.calculate (mbt_reward_eval)."""
    from mbt_gym_b200.rewards.RewardFunctions import CjMmCriterion, PnL, RunningInventoryPenalty

    cur = np.array([[120, 2, 0.5, 100.0]]); act = np.array([[1, 1.0]]); nxt = np.array([[20, 3, 0.7, 100.05]])
    expected = (nxt[:, 0] + nxt[:, 1] * nxt[:, 3]) - (cur[:, 0] + cur[:, 1] * cur[:, 3])
    assert PnL().calculate(cur, act, nxt)[0] == expected[0]
    rip = RunningInventoryPenalty(0.01, 1)
    assert rip.calculate(cur, act, nxt)[0] == pytest.approx(expected[0] - 0.01 * 0.2 * 9, abs=1e-5)
    step = 0.2
    obs = [np.array([[100.0, 0, 0.0, 100]]), np.array([[0.5, 1, step, 101]]), np.array([[102.0, 0, 2 * step, 102]]),
           np.array([[103.0, 0, 3 * step, 103]]), np.array([[206.5, -1, 4 * step, 104]]), np.array([[103.0, 0, 5 * step, 103]])]
    acts = [np.array([[0.5, 0.5]]), np.array([[0.5, 1]]), np.array([[0.5, 0.5]]), np.array([[1, 0.5]]), np.array([[0.5, 0.5]])]
    mm = CjMmCriterion(0.01, 1, terminal_time=1.0)
    mm.reset(obs[0])
    tot_mm = sum(float(mm.calculate(obs[i], acts[i], obs[i + 1], obs[i + 1][:, 2] == 1)[0]) for i in range(5))
    tot_rip = sum(float(rip.calculate(obs[i], acts[i], obs[i + 1], obs[i + 1][:, 2] == 1)[0]) for i in range(5))
    assert tot_mm == pytest.approx(tot_rip, abs=1e-5)


def test_normalise_rewards_bootstrap_and_setters():
    spec = dict(SPECS["as_pnl"], N=256)
    env = build_facade_env(spec)
    env.normalise_rewards_ = True
    env.reward_scaling = 1 / env._get_inventory_neutral_rewards(num_total_trajectories=20000)
    # inventory-neutral fixed action 1/kappa: expected PnL per episode = 2 * lambda * T * exp(-1) / kappa
    expect = 2 * 140 * np.exp(-1) / 1.5
    assert abs(1 / env.reward_scaling - expect) < 1.5
    env.num_trajectories = 128            # setter rebuilds the device handle lazily
    obs = env.reset()
    assert obs.shape == (128, 4)
    _, rew, _, infos = env.step(np.full((128, 2), 0.7))
    assert rew.shape == (128,) and len(infos) == 128
    env.close()


def test_wrappers_and_full_size_step_properties():
    """(1) the reference's gym wrappers on top of the device env; (2) size-independent properties of the STEP path at
    BASELINE size N = 2^20: PnL rewards telescope to the change in marked-to-market wealth, inventories stay integral,
    the empirical fill rate matches lambda*dt*exp(-kappa*delta), time advances by dt."""
    from mbt_gym_b200.gym.wrappers import ReduceStateSizeWrapper

    spec = dict(SPECS["as_pnl"], N=1 << 20)
    env = build_facade_env(spec)
    wrapped = ReduceStateSizeWrapper(env)
    assert wrapped.observation_space.shape == (2,)
    obs = wrapped.reset()
    assert obs.shape == (1 << 20, 2) and np.all(obs == 0)
    a = np.full((1 << 20, 2), 0.7)
    total = np.zeros(1 << 20)
    steps = 25
    for k in range(steps):
        o, r, d, _ = wrapped.step(a)
        total += r
    state = env.state
    wealth = state[:, 0] + state[:, 1] * state[:, 3]
    np.testing.assert_allclose(total, wealth - 0.0, rtol=0, atol=1e-9)          # sum of PnL rewards telescopes
    assert np.all(state[:, 1] == np.round(state[:, 1]))                         # inventory moves in whole units
    assert state[0, 2] == pytest.approx(steps * env.step_size)
    # each side fills with probability 0.7 * exp(-1.05) per step; |q| changes only through fills
    p = 0.7 * np.exp(-1.5 * 0.7)
    expected_var_q = 2 * steps * p * (1 - p)                                    # bid/ask fills independent Bernoulli(p)
    assert abs(state[:, 1].var() - expected_var_q) < 6 * expected_var_q * np.sqrt(2 / (1 << 20))
    assert abs(state[:, 1].mean()) < 6 * np.sqrt(expected_var_q / (1 << 20))
    assert abs(state[:, 3].var() - 4.0 * steps * env.step_size) < 6 * 4.0 * steps * env.step_size * np.sqrt(2 / (1 << 20))
    env.close()


def test_checkpoint_resume_is_exact():
    g = Golden("cjmm")   # random initial inventories + late start: exercises q0 column, t0, counters
    env = build_facade_env(SPECS["cjmm"])
    env.reset()
    for k in range(10):
        env.step(g.actions[k])
    blob = env.save_checkpoint()
    tail = [tuple(np.array(x, copy=True) for x in env.step(g.actions[k])[:2]) for k in range(10, 30)]
    other = build_facade_env(dict(SPECS["cjmm"], seed=999))      # a different key: the checkpoint must carry its own
    other.reset()
    other.load_checkpoint(blob)
    for k, (o, r) in zip(range(10, 30), tail):
        o2, r2, _, _ = other.step(g.actions[k])
        assert_same(o2, o, what=f"resumed obs {k}"); assert_same(r2, r, what=f"resumed rew {k}")
        assert_same(o2, g.obs[k], what=f"fixture obs {k}")
    with pytest.raises(Exception):
        build_facade_env(dict(SPECS["cjmm"], N=7)).load_checkpoint(blob)
    env.close(); other.close()


def test_generate_trajectory_fused_equals_stepwise_generate_trajectory():
    """The fused, recording rollout returns exactly what the reference-style step-by-step helper returns."""
    from mbt_gym_b200.agents.BaselineAgents import AvellanedaStoikovAgent, CarteaJaimungalMmAgent, CarteaJaimungalOeAgent
    from mbt_gym_b200.gym.helpers.generate_trajectory import generate_trajectory, generate_trajectory_fused

    import warnings
    for name, make_agent in [("as_pnl", lambda e: AvellanedaStoikovAgent(0.1, e)),
                             ("cjmm", lambda e: CarteaJaimungalMmAgent(e)),
                             ("oe_ou_cjoe", lambda e: CarteaJaimungalOeAgent(env=e)),
                             ("hawkes_normalised", None)]:
        spec = dict(SPECS[name], N=300)
        if name == "cjmm":
            spec.update(initial_inventory=0, start_time=0.0)
        if name == "hawkes_normalised":
            spec.update(normalise_action=False)
        env = build_facade_env(spec)
        if make_agent is None:
            from mbt_gym_b200.agents.BaselineAgents import FixedSpreadAgent
            agent = FixedSpreadAgent(env, half_spread=0.8, offset=0.1)
        else:
            agent = make_agent(env)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            env.seed(123)
            o1, a1, r1 = generate_trajectory(env, agent)
            env.seed(123)
            o2, a2, r2 = generate_trajectory_fused(env, agent)
        assert o2.shape == o1.shape and a2.shape == a1.shape and r2.shape == r1.shape, name
        assert_same(a2, a1, what=f"{name} actions")
        assert_same(o2, o1, what=f"{name} observations")
        assert_same(r2, r1, what=f"{name} rewards")
        env.close()


@pytest.mark.parametrize("name,cols", [("as_pnl", [1, 2]), ("as_pnl_normalised", [1, 2]), ("hawkes_pnl", [1, 3, 5]),
                                       ("oe_ou_cjoe", [0, 4]), ("oe_temp_transient", [1, 2, 3, 4])])
@pytest.mark.parametrize("precision", ["float64", "float32"])
def test_fused_column_select_equals_host_side_wrapper(name, cols, precision):
    """`ReduceStateSizeWrapper` fused into the kernel's observation store == the reference behaviour (select on the host)."""
    from mbt_gym_b200.gym.wrappers import ReduceStateSizeWrapper

    g = Golden(name)
    fused = ReduceStateSizeWrapper(build_facade_env(SPECS[name], precision=precision), cols)
    plain = ReduceStateSizeWrapper(build_facade_env(SPECS[name], precision=precision), cols, fuse=False)
    assert fused._fused and not plain._fused and fused.observation_space.shape == (len(cols),)
    assert fused.env._native is None or True
    o1, o2 = fused.reset(), plain.reset()
    assert o1.shape == (g.cfg.num_trajectories, len(cols))
    assert_same(o1, o2, what=f"{name} reset")
    for k in range(12):
        a = g.actions[k]
        o1, r1, d1, _ = fused.step(a)
        o2, r2, d2, _ = plain.step(a)
        assert_same(o1, o2, what=f"{name} obs {k}"); assert_same(r1, r2, what=f"{name} rew {k}")
        if precision == "float64":
            assert_same(o1, g.obs[k][:, cols], exact=g.exact, what=f"{name} vs fixture {k}")
    assert fused.env.state.shape[1] == g.obs.shape[2], "env.state keeps every column"
    fused.env.close(); plain.env.close()


@pytest.mark.parametrize("name", ["as_pnl", "hawkes_normalised", "oe_ou_cjoe", "cjmm", "power_fill", "triangular_fill"])
def test_float32_io_over_float64_arithmetic_is_the_rounded_float64_result(name):
    """io_dtype=float32: actions arrive as float32, the dynamics run in float64 exactly as before, observations and
    rewards leave rounded to float32 -- so they must equal np.float32(<float64 run fed the same float32 actions>)."""
    g = Golden(name)
    e32 = build_facade_env(SPECS[name], io_dtype=np.float32)
    e64 = build_facade_env(SPECS[name])
    o32, o64 = e32.reset(), e64.reset()
    assert o32.dtype == np.float32 and o64.dtype == np.float64
    assert_same(o32, o64.astype(np.float32), what=f"{name} reset")
    for k in range(15):
        a32 = g.actions[k].astype(np.float32)
        o32, r32, d32, _ = e32.step(a32)
        o64, r64, d64, _ = e64.step(a32.astype(np.float64))
        assert o32.dtype == np.float32 and r32.dtype == np.float32
        assert_same(o32, o64.astype(np.float32), what=f"{name} obs {k}")
        assert_same(r32, r64.astype(np.float32), what=f"{name} rew {k}")
    assert_same(e32.state, e64.state, what=f"{name} state stays float64 and identical")
    e32.close(); e64.close()


def test_cuda_graph_replay_of_captured_episodes_equals_eager_episodes():
    """Whole episodes (reset + a torch policy + step, all on the device) captured into ONE CUDA graph: thanks to the
    device-resident counter base (mbt_fold_counters) every replay is a new episode with fresh random numbers, and R
    replays are bit-identical to R episodes stepped eagerly from the same seed; eager calls afterwards continue the
    same random streams."""
    import torch

    # random initial inventories: the reset draws random numbers too
    spec = dict(golden_specs()["cjmm"], N=5000, n_steps=12, seed=99, start_time=0.0)
    sign = torch.tensor([[1.0, -1.0]], dtype=torch.float64, device="cuda")

    def policy(obs):  # any torch function of the observation; (N, 2) depths
        return 0.6 + 0.2 * torch.tanh(obs[:, 1:2]) * sign

    def eager_episode(env):
        obs = env.reset_device()
        ret = torch.zeros(spec["N"], dtype=torch.float64, device="cuda")
        for _ in range(spec["n_steps"]):
            obs, rew, dones, _ = env.step(policy(obs))
            ret = ret + rew
        assert bool(dones[0])
        torch.cuda.synchronize()
        return obs.cpu().numpy(), ret.cpu().numpy()

    eager_env = build_facade_env(spec)
    want = [eager_episode(eager_env) for _ in range(5)]
    eager_env.close()
    assert not np.array_equal(want[0][1], want[1][1])

    env = build_facade_env(spec)
    s = torch.cuda.Stream()
    obs_t = torch.empty((spec["N"], 4), dtype=torch.float64, device="cuda")
    ret_t = torch.zeros(spec["N"], dtype=torch.float64, device="cuda")
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):  # binds the handle to the capture stream and warms torch up; consumes no random numbers
        env._ensure_native().set_stream(s.cuda_stream)
        policy(obs_t.zero_())
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        env.reset_device(out=obs_t)
        ret_t.zero_()
        for _ in range(spec["n_steps"]):
            o, rew, _d, _ = env.step(policy(obs_t))
            obs_t.copy_(o)
            ret_t += rew
        env.fold_counters()
    for r in range(3):
        g.replay()
        torch.cuda.synchronize()
        assert_same(obs_t.cpu().numpy(), want[r][0], what=f"replay {r}: final observations")
        assert_same(ret_t.cpu().numpy(), want[r][1], what=f"replay {r}: episode returns")
    assert env._native.clock()["n_step"] == 3 * spec["n_steps"] and env._native.clock()["n_episode"] == 3
    # eager continuation after the replays: episode 4 of the same streams; checkpoint / resume carries the counters
    blob = env.save_checkpoint()
    with torch.cuda.stream(s):
        got = eager_episode(env)
    assert_same(got[0], want[3][0], what="eager episode after replays: observations")
    assert_same(got[1], want[3][1], what="eager episode after replays: returns")
    env.load_checkpoint(blob)
    with torch.cuda.stream(s):
        got = eager_episode(env)
    assert_same(got[1], want[3][1], what="episode after checkpoint restore")
    with torch.cuda.stream(s):
        got = eager_episode(env)
    assert_same(got[1], want[4][1], what="next episode after checkpoint restore")
    env.close()


def test_vecenv_adapter_survives_the_access_patterns_of_sb3_collect_rollouts_and_vecmonitor():
    """stable-baselines3 is not installed here, so its two consumers of `infos` / `dones` are restated from its source
    (common/on_policy_algorithm.py collect_rollouts + base_class._update_info_buffer, common/vec_env/vec_monitor.py
    step_wait) and run against the adapter for two whole episodes: indexing, slicing, copying, item assignment and
    iteration must all work on ordinary steps (the env's list of empty dicts) and on episode ends (_TerminalInfos)."""
    from mbt_gym_b200.gym.StableBaselinesTradingEnvironment import StableBaselinesTradingEnvironment

    spec = dict(SPECS["as_pnl"], N=257, n_steps=12)
    venv = StableBaselinesTradingEnvironment(build_facade_env(spec))
    n = venv.num_envs
    obs = venv.reset()
    ep_returns, ep_lengths = np.zeros(n), np.zeros(n, dtype=int)   # VecMonitor state
    ep_info_buffer, terminal_seen = [], 0
    rng = np.random.default_rng(0)
    for t in range(24):
        actions = rng.uniform(0.1, 1.2, size=(n, 2)).astype(np.float32)   # SB3 policies emit float32
        venv.step_async(actions)
        new_obs, rewards, dones, infos = venv.step_wait()
        # --- VecMonitor.step_wait
        ep_returns += rewards
        ep_lengths += 1
        new_infos = list(infos[:])
        for i in range(len(dones)):
            if dones[i]:
                info = infos[i].copy()
                info["episode"] = {"r": ep_returns[i], "l": ep_lengths[i], "t": 0.0}
                ep_returns[i] = 0
                ep_lengths[i] = 0
                new_infos[i] = info
        infos = new_infos
        # --- BaseAlgorithm._update_info_buffer
        for idx, info in enumerate(infos):
            maybe_ep_info = info.get("episode")
            if maybe_ep_info is not None:
                ep_info_buffer.extend([maybe_ep_info])
        # --- OnPolicyAlgorithm.collect_rollouts: bootstrap with the terminal observation on time-limit truncation
        for idx, done in enumerate(dones):
            if done and infos[idx].get("terminal_observation") is not None:
                terminal_seen += 1
                assert infos[idx]["terminal_observation"].shape == obs.shape[1:]
                assert not infos[idx].get("TimeLimit.truncated", False)
        assert new_obs.shape == obs.shape and rewards.shape == (n,) and dones.shape == (n,) and dones.dtype == bool
        obs = new_obs
    assert terminal_seen == 2 * n and len(ep_info_buffer) == 2 * n
    assert all(e["l"] == 12 for e in ep_info_buffer)
    assert venv.env_is_wrapped(object) == [False] * n
    venv.close()
