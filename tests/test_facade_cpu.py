"""Host-side mirror of the reference interface (no GPU needed): spaces, bounds, and the flattened mbt_config.

For every fixture the facade built with the reference's keyword arguments must flatten to EXACTLY the config bytes that
tools/make_golden.py read off the constructed reference environment (spaces, normalisation constants, max_cash, ...).
"""
import numpy as np
import pytest

from mbt_gym_b200 import _abi
from tests.helpers import Golden, build_facade_env, golden_names, golden_specs

SPECS = golden_specs()


@pytest.mark.parametrize("name", golden_names())
def test_facade_flattens_to_the_reference_config(name):
    g = Golden(name)
    env = build_facade_env(SPECS[name])
    cfg = env._build_config()
    want = g.config(_abi.MBT_F64)
    for field, _t in _abi.mbt_config._fields_:
        a, b = getattr(cfg, field), getattr(want, field)
        if hasattr(a, "__len__"):
            a, b = list(a), list(b)
        assert a == b, f"{name}: mbt_config.{field}: facade {a} != reference {b}"
    assert bytes(cfg) == bytes(want)


def test_default_constructor_spaces_match_reference_probe():
    """default ctor -> obs Box [-1,1]^4, action Box [-1,1]^2; original bounds [-21600,-10000,0,92]..."""
    from mbt_gym_b200.gym.TradingEnvironment import TradingEnvironment

    env = TradingEnvironment(num_trajectories=3)
    assert env.observation_space.shape == (4,) and env.action_space.shape == (2,)
    np.testing.assert_array_equal(env.original_observation_space.low, np.float32([-21600, -10000, 0, 92]))
    np.testing.assert_array_equal(env.original_observation_space.high, np.float32([21600, 10000, 1, 108]))
    np.testing.assert_allclose(env.original_action_space.high, np.float32(-np.log(0.01) / 1.5))
    assert env.step_size == 1.0 / 200 and env.num_trajectories == 3
    # setters propagate to the processes (TradingEnvironment.py:154-178)
    env.num_trajectories = 5
    assert all(p.num_trajectories == 5 for p in env.stochastic_processes.values())
    env.step_size = 0.01
    assert all(p.step_size == 0.01 for p in env.stochastic_processes.values())
    # normalisation helpers are the reference's affine maps
    obs = np.array([[0.0, 0.0, 0.005, 100.0]])
    n = env.normalise_observation(obs)
    np.testing.assert_allclose(n, [[0.0, 0.0, -0.99, 0.0]], atol=1e-12)
    np.testing.assert_allclose(env.normalise_observation(n, inverse=True), obs)
    a = np.array([[0.5, -0.5]])
    np.testing.assert_allclose(env.normalise_action(env.normalise_action(a, inverse=True)), a)


def test_agents_host_policies_match_reference_formulas():
    from mbt_gym_b200.agents.BaselineAgents import (AvellanedaStoikovAgent, CarteaJaimungalMmAgent,
                                                     CarteaJaimungalOeAgent, FixedSpreadAgent)

    env = build_facade_env(SPECS["as_pnl"])
    ag = AvellanedaStoikovAgent(risk_aversion=0.1, env=env)
    st = np.array([[0.0, 2.0, 0.25, 100.0], [0.0, -1.0, 0.25, 100.0]])
    act = ag.get_action(st)
    gam, sig, kap, tau = 0.1, 2.0, 1.5, 0.75
    spread = gam * sig ** 2 * tau + 2 / gam * np.log(1 + gam / kap)
    np.testing.assert_allclose(act[0], [2 * gam * sig ** 2 * tau + spread / 2, -2 * gam * sig ** 2 * tau + spread / 2])
    pol = ag.to_policy(env)
    assert pol.kind == _abi.MBT_POL_AVELLANEDA_STOIKOV and pol.as_sigma_sq == 4.0
    assert FixedSpreadAgent(env, 1.0, 0.25).get_action(st).tolist() == [[0.75, 1.25]] * env.num_trajectories

    cj_env = build_facade_env(SPECS["cjmm"])
    cj = CarteaJaimungalMmAgent(env=cj_env)
    state = np.zeros((cj_env.num_trajectories, 4))
    state[:, 1] = np.arange(cj_env.num_trajectories) % 7 - 3
    d = cj.get_action(state)
    assert d.shape == (cj_env.num_trajectories, 2) and np.all(np.isfinite(d))
    # closed-form value function of the CJP-2015 notebook (Test_2 set 1): h(0, 0) with phi=0.01, alpha=0.001, Q=100
    v = cj.calculate_true_value_function(np.array([[0.0, 0.0, 0.0, 100.0]] * cj_env.num_trajectories))
    assert abs(v[0, 0] - 68.25583476) < 1e-6 if v.ndim == 2 else abs(v[0] - 68.25583476) < 1e-6

    oe_env = build_facade_env(SPECS["oe_ou_cjoe"])
    oe = CarteaJaimungalOeAgent(env=oe_env)
    a = oe.get_action(np.zeros((oe_env.num_trajectories, 5)))
    assert a.shape == (oe_env.num_trajectories, 1) and np.all(a == a[0, 0]) and np.isfinite(a[0, 0])


def test_unsupported_models_fail_loudly():
    from mbt_gym_b200.stochastic_processes.StochasticProcessModel import StochasticProcessModel

    class MyProcess(StochasticProcessModel):
        def __init__(self):
            super().__init__([[0]], [[1]], 0.1, 1.0, [[0]], 1)

    with pytest.raises(NotImplementedError):
        MyProcess()._flatten(_abi.new_config())
    with pytest.raises(NotImplementedError):
        MyProcess().update(None, None, None)


def test_sb_agent_adapter_duck_types_a_stable_baselines_model():
    """SbAgent (reference agents/SbAgent.py:8-26) needs only predict / action_space / env / learn from the model."""
    from mbt_gym_b200.agents.SbAgent import SbAgent

    class Space:
        shape = (2,)

    class Env:
        num_trajectories = 5

    class Model:
        action_space, env, learned = Space(), Env(), 0

        def predict(self, obs, deterministic=False):
            assert deterministic and obs.shape == (5, 2)
            return np.stack([obs[:, 0] + 1, obs[:, 1] * 2], axis=1).reshape(-1), None  # flat, like some SB3 policies

        def learn(self, total_timesteps):
            self.learned += total_timesteps

    model = Model()
    agent = SbAgent(model, reduced_training_indices=[1, 2])
    state = np.arange(20.0).reshape(5, 4)
    act = agent.get_action(state)
    assert act.shape == (5, 2)
    np.testing.assert_array_equal(act, np.stack([state[:, 1] + 1, state[:, 2] * 2], axis=1))
    agent.train(123)
    assert model.learned == 123 and agent.num_trajectories == 5


def test_terminal_infos_is_a_lazy_but_real_list():
    """SB3's VecMonitor / collect_rollouts index, slice, copy and assign into `infos`: the episode-end infos must be a list."""
    import copy
    import pickle

    from mbt_gym_b200.gym.StableBaselinesTradingEnvironment import _TerminalInfos

    obs = np.arange(12.0).reshape(4, 3)
    infos = _TerminalInfos(obs, (np.array([1.0, 2, 3, 4]), 5, 0.1))
    assert isinstance(infos, list) and len(infos) == 4 and bool(infos)
    assert infos[2]["episode"] == {"r": 3.0, "l": 5, "t": 0.1} and np.array_equal(infos[-1]["terminal_observation"], obs[3])
    infos[1]["TimeLimit.truncated"] = True      # mutation of a lazily built item sticks
    infos[3] = {"replaced": 1}                  # item assignment before anything materialised the list
    new_infos = list(infos[:])                  # what VecMonitor does
    assert new_infos[1]["TimeLimit.truncated"] is True and new_infos[3] == {"replaced": 1} and len(new_infos) == 4
    assert [i.get("episode", {}).get("r") for i in infos] == [1.0, 2.0, 3.0, None]
    infos.append({})
    assert len(infos) == 5 and infos == new_infos + [{}]
    again = _TerminalInfos(obs)
    assert pickle.loads(pickle.dumps(again))[2]["terminal_observation"][1] == 7.0
    assert isinstance(copy.copy(_TerminalInfos(obs)), list) and len(copy.deepcopy(_TerminalInfos(obs))) == 4
    with pytest.raises(IndexError):
        _TerminalInfos(obs)[4]


def test_pinned_pool_recycles_only_unreferenced_blocks(monkeypatch):
    """The output pool's recycling rule, on ordinary memory (no CUDA here): a block is reused only when no array or view of
    it is alive."""
    import ctypes as C

    from mbt_gym_b200 import _lib

    class HostBlock(_lib._PinnedBlock):
        def __init__(self, nbytes, device=0):
            self.nbytes = max(int(nbytes), 1)
            self._buf = (C.c_char * self.nbytes)()
            self._ptr = C.c_void_p(C.addressof(self._buf))

        def __del__(self):
            pass

    monkeypatch.setattr(_lib, "_PinnedBlock", HostBlock)
    pool = _lib.PinnedPool(max_bytes=5 * 96)
    a, b = pool.get((4, 3), np.float64), pool.get((4, 3), np.float64)
    assert a.ctypes.data != b.ctypes.data and len(pool._blocks) == 2
    a[:] = 1.0
    view = a[:, 1]
    del a
    c = pool.get((4, 3), np.float64)            # `view` still references the first block: a third one is made
    assert len(pool._blocks) == 3 and np.all(view == 1.0)
    c[:] = 2.0
    assert np.all(view == 1.0)
    del b, c, view
    d = pool.get((4, 3), np.float64)
    assert len(pool._blocks) == 3               # reuse
    e, f = pool.get((4, 3), np.float64), pool.get((4, 3), np.float64)
    g = pool.get((4, 3), np.float64), pool.get((4, 3), np.float64)
    assert len(pool._blocks) == 5 and g[1] is not None
    assert pool.get((4, 3), np.float64) is None, "beyond max_bytes the pool declines (the caller falls back to np.empty)"
    del d, e, f, g


def test_public_attribute_edits_bump_the_version_private_ones_do_not():
    from mbt_gym_b200 import _track

    env = build_facade_env(SPECS["as_pnl"])
    v0 = _track.version[0]
    env._scratch = 1
    env.model_dynamics.midprice_model._scratch = 2
    assert _track.version[0] == v0
    env.max_inventory = 7
    env.model_dynamics.midprice_model.volatility = 3.0
    env.reward_function.anything = 1
    assert _track.version[0] == v0 + 3
    assert env._build_config().max_inventory == 7.0 and env._build_config().mid_vol == 3.0
