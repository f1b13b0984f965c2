"""Checks that need the UNMODIFIED reference at /root/reference (build container only; skipped on the GPU box).

  * the oracle equals the reference step-for-step under injected draws (the generator of tests/golden, re-run live);
  * the committed fixtures are what the reference produces today;
  * the facade's host agents return exactly what the reference's agents return;
  * the NumPy port used as CPU baseline reproduces the reference's notebook golden (Test_1, seed 50).
"""
import numpy as np
import pytest

from oracle import ref_shim as R
from tests.helpers import Golden, build_facade_env, golden_specs

pytestmark = pytest.mark.reference
SPECS = golden_specs()


@pytest.mark.parametrize("name", ["as_pnl_tight", "cjmm", "hawkes_normalised", "oe_ou_cjoe", "limit_and_market", "two_episodes"])
def test_oracle_equals_live_reference_and_fixture_is_current(name):
    spec = SPECS[name]
    out = R.run_pair(spec, n_episodes=spec.get("n_episodes", 1))
    assert np.array_equal(out["ref_obs"], out["orc_obs"]) and np.array_equal(out["ref_rew"], out["orc_rew"])
    assert out["ref_done"] == out["orc_done"] and np.array_equal(out["ref_reset"], out["orc_reset"])
    g = Golden(name)
    assert np.array_equal(out["ref_obs"], g.obs) and np.array_equal(out["ref_rew"], g.rew)
    assert bytes(out["cfg"]) == bytes(g.cfg)


def test_host_agents_equal_reference_agents():
    R.import_reference()
    from mbt_gym.agents import BaselineAgents as RefAgents

    from mbt_gym_b200.agents import BaselineAgents as Agents

    rng = np.random.default_rng(0)
    for name, make_ref, make_mine in [
        ("as_pnl", lambda e: RefAgents.AvellanedaStoikovAgent(0.1, e), lambda e: Agents.AvellanedaStoikovAgent(0.1, e)),
        ("cjmm", lambda e: RefAgents.CarteaJaimungalMmAgent(e), lambda e: Agents.CarteaJaimungalMmAgent(e)),
        ("oe_ou_cjoe", lambda e: RefAgents.CarteaJaimungalOeAgent(env=e), lambda e: Agents.CarteaJaimungalOeAgent(env=e)),
    ]:
        spec = SPECS[name]
        ref_env, my_env = R.build_reference_env(spec), build_facade_env(spec)
        ref_agent, my_agent = make_ref(ref_env), make_mine(my_env)
        n, d = spec["N"], ref_env.observation_space.shape[0]
        for t in (0.0, 0.25, 0.995):
            state = np.zeros((n, d))
            state[:, 1] = rng.integers(-12, 13, size=n)
            state[:, 2] = t
            state[:, 3] = 100.0
            np.testing.assert_array_equal(my_agent.get_action(state), ref_agent.get_action(state), err_msg=f"{name} t={t}")
        if name == "cjmm":
            np.testing.assert_array_equal(np.asarray(my_agent.calculate_true_value_function(state)).reshape(-1),
                                          np.asarray(ref_agent.calculate_true_value_function(state)).reshape(-1))


def test_numpy_port_reproduces_notebook_golden():
    """Test_1 notebook, gamma=0.1 (notebooks/Test_1_-_replicate_AS_original_results.ipynb:219-231): second rollout."""
    from oracle import numpy_port as P

    env = P.NumpyPortEnv("as", N=1000, seed=50)
    for _ in range(2):
        obs = env.reset()
        R_ = np.zeros(1000)
        acts = []
        while True:
            a = P.as_agent_action(obs, 0.1, 2.0, 1.5, 1.0)
            acts.append(a)
            obs, r, d, _ = env.step(a)
            R_ += r
            if d[0]:
                break
    got = [2 * np.mean(acts), R_.mean(), R_.std(), obs[:, 1].mean(), obs[:, 1].std()]
    np.testing.assert_allclose(got, [1.49177, 64.872139, 6.692567, 0.201, 2.893544], rtol=0, atol=6e-7)


def _space_arrays(space):
    return [np.asarray(getattr(space, k)) for k in ("low", "high") if hasattr(space, k)]


def test_spaces_and_host_attributes_equal_the_reference_for_every_configuration():
    """What an SB3 policy sees before the first step: observation / action spaces (and their un-normalised originals),
    max_cash, step size, start time, `initial_state` -- equal to the reference's for every fixture specification and for
    the randomly drawn configurations (all model kinds incl. Heston, Hawkes, the four fill models, the four dynamics)."""
    from tests.test_random_configs import random_specs

    specs = list(SPECS.items()) + [(f"random{i}", s) for i, s in enumerate(random_specs(40, 20260925))]
    for name, spec in specs:
        ref, mine = R.build_reference_env(spec), build_facade_env(spec)
        for attr in ("observation_space", "action_space", "original_observation_space", "original_action_space"):
            if not hasattr(ref, attr):
                assert not hasattr(mine, attr) or attr.startswith("original"), (name, attr)
                continue
            a, b = _space_arrays(getattr(ref, attr)), _space_arrays(getattr(mine, attr))
            assert len(a) == len(b), (name, attr)
            for x, y in zip(a, b):
                assert x.shape == y.shape and x.dtype == y.dtype, (name, attr, x.shape, y.shape, x.dtype, y.dtype)
                np.testing.assert_array_equal(x, y, err_msg=f"{name} {attr}")
        assert mine.max_cash == ref.max_cash and mine.max_inventory == ref.max_inventory, name
        assert mine.step_size == ref.step_size and mine.n_steps == ref.n_steps and mine.terminal_time == ref.terminal_time
        assert mine._get_start_time() == ref._get_start_time(), name
        assert mine.num_trajectories == ref.num_trajectories
        if not isinstance(spec.get("initial_inventory", 0), list):  # (random inventories come from different generators)
            np.testing.assert_array_equal(mine.initial_state, ref.initial_state, err_msg=f"{name} initial_state")


def test_wrappers_equal_the_reference_wrappers_on_the_same_env_outputs():
    """gym/wrappers.py is host-side glue: fed the same environment outputs, the facade's wrappers must return exactly
    what the reference's return (including NormaliseASObservation's asymmetric reset / step maps)."""
    R.import_reference()
    import gym  # the stub under oracle/gym_stub
    from mbt_gym.gym import wrappers as RefW

    from mbt_gym_b200 import spaces as my_spaces
    from mbt_gym_b200.gym import wrappers as MyW

    rng = np.random.default_rng(5)
    low, high = np.float32([-2160.0, -20.0, 0.0, 92.0]), np.float32([2160.0, 20.0, 1.0, 108.0])
    frames = [rng.uniform(low, high, size=(7, 4)) for _ in range(4)]
    rewards = [rng.normal(size=7) for _ in range(4)]

    class Reward:
        per_step_inventory_aversion, terminal_inventory_aversion = 0.01, 0.5

    def make_env(box_cls):
        class DummyEnv:
            observation_space = box_cls(low=low, high=high)
            reward_function = Reward()
            spec = None

            def __init__(self):
                self.k = 0

            def reset(self):
                self.k = 0
                return frames[0].copy()

            def step(self, action):
                self.k += 1
                done = self.k == 3
                return frames[self.k].copy(), rewards[self.k].copy(), (done if box_cls is gym.spaces.box.Box else np.full(7, done)), {}

        return DummyEnv()

    for ref_cls, my_cls, kwargs in [(RefW.ReduceStateSizeWrapper, MyW.ReduceStateSizeWrapper, {}),
                                    (RefW.ReduceStateSizeWrapper, MyW.ReduceStateSizeWrapper, {"list_of_state_indices": [3, 0]}),
                                    (RefW.NormaliseASObservation, MyW.NormaliseASObservation, {}),
                                    (RefW.RemoveTerminalRewards, MyW.RemoveTerminalRewards, {})]:
        ref, mine = ref_cls(make_env(gym.spaces.box.Box), **kwargs), my_cls(make_env(my_spaces.Box), **kwargs)
        if hasattr(ref, "observation_space") and ref_cls is not RefW.RemoveTerminalRewards:
            np.testing.assert_array_equal(ref.observation_space.low, mine.observation_space.low)
            np.testing.assert_array_equal(ref.observation_space.high, mine.observation_space.high)
        np.testing.assert_array_equal(ref.reset(), mine.reset())
        for _ in range(3):
            (o1, r1, d1, _i1), (o2, r2, d2, _i2) = ref.step(None), mine.step(None)
            np.testing.assert_array_equal(o1, o2, err_msg=ref_cls.__name__)
            np.testing.assert_array_equal(r1, r2, err_msg=ref_cls.__name__)
            assert bool(np.asarray(d1).reshape(-1)[0]) == bool(np.asarray(d2).reshape(-1)[0])


def test_generate_trajectory_helper_equals_the_reference_helper_on_the_same_env():
    """gym/helpers/generate_trajectory.py (the canonical rollout loop every notebook uses) is host-side glue: on the same
    environment outputs and the same agent it must record exactly what the reference's helper records -- with and
    without log-probabilities, for several trajectories and for one."""
    import torch

    R.import_reference()
    from mbt_gym.gym.helpers.generate_trajectory import generate_trajectory as ref_generate

    from mbt_gym_b200.gym.helpers.generate_trajectory import generate_trajectory as my_generate

    class Space:
        def __init__(self, d):
            self.shape = (d,)

    class DummyEnv:
        n_steps = 6

        def __init__(self, n):
            self.num_trajectories, self.observation_space, self.action_space = n, Space(4), Space(2)
            self.rng = np.random.default_rng(11)

        def seed(self, seed=None):
            self.rng = np.random.default_rng(seed)

        def reset(self):
            self.k = 0
            return self.rng.normal(size=(self.num_trajectories, 4))

        def step(self, action):
            self.k += 1
            obs = self.rng.normal(size=(self.num_trajectories, 4)) + action.sum()
            rew = self.rng.normal(size=(self.num_trajectories,))
            done = self.k == self.n_steps
            return obs, rew, (np.full(self.num_trajectories, done) if self.num_trajectories > 1 else done), {}

    class DummyAgent:
        def get_action(self, obs, include_log_probs=False):
            a = np.stack([obs[:, 0] * 0.5, obs[:, 1] - 1.0], axis=1)
            return (a, torch.from_numpy(a * 0.1)) if include_log_probs else a

    for n in (5, 1):
        for with_lp in (False, True):
            ref = ref_generate(DummyEnv(n), DummyAgent(), seed=3, include_log_probs=with_lp)
            mine = my_generate(DummyEnv(n), DummyAgent(), seed=3, include_log_probs=with_lp)
            assert len(ref) == len(mine) == (4 if with_lp else 3)
            for a, b in zip(ref, mine):
                np.testing.assert_array_equal(np.asarray(a), np.asarray(b))
