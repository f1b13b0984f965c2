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
