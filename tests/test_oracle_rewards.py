"""The reference's own unit tests (mbt_gym/rewards/tests/testRewardFunctions.py:33-135), restated on the oracle's
reward functions: PnL, RunningInventoryPenalty, and the CjMm telescoping identity in its three variants."""
from copy import deepcopy

import numpy as np
import pytest

from mbt_gym_b200 import _abi
from oracle import oracle as O

STEP_SIZE = 0.2
TEST_CURRENT_STATE = np.array([[120, 2, 0.5, 100.0]])
TEST_ACTION = np.array([[1, 1.0]])
TEST_NEXT_STATE = np.array([[20, 3, 0.5 + STEP_SIZE, 100.05]])  # buy order gets filled
TERMINAL_TIME = 1.0
MOCK_OBSERVATIONS = [np.array([[100.0, 0, 0.0, 100]]), np.array([[0.5, 1, STEP_SIZE, 101]]),
                     np.array([[102.0, 0, 2 * STEP_SIZE, 102]]), np.array([[103.0, 0, 3 * STEP_SIZE, 103]]),
                     np.array([[206.5, -1, 4 * STEP_SIZE, 104]]), np.array([[103.0, 0, 5 * STEP_SIZE, 103]])]
MOCK_ACTIONS = [np.array([[0.5, 0.5]]), np.array([[0.5, 1]]), np.array([[0.5, 0.5]]), np.array([[1, 0.5]]),
                np.array([[0.5, 0.5]])]
PHI, ALPHA = 0.01, 1


def cfg_for(reward, phi=0.0, alpha=0.0):
    return _abi.new_config(precision=_abi.MBT_F64, num_trajectories=1, n_steps=5, dynamics=_abi.MBT_DYN_LIMIT,
                           midprice=_abi.MBT_MID_BM, arrival=_abi.MBT_ARR_POISSON, fill=_abi.MBT_FILL_EXPONENTIAL,
                           reward=reward, rew_phi=phi, rew_alpha=alpha, rew_exponent=2.0, terminal_time=TERMINAL_TIME,
                           step_size=STEP_SIZE, rew_terminal_time=TERMINAL_TIME)


def test_pnl_per_step_reward():
    expected = (TEST_NEXT_STATE[:, 0] + TEST_NEXT_STATE[:, 1] * TEST_NEXT_STATE[:, 3]) - (
        TEST_CURRENT_STATE[:, 0] + TEST_CURRENT_STATE[:, 1] * TEST_CURRENT_STATE[:, 3])
    actual = O.reward_eval(cfg_for(_abi.MBT_REW_PNL), TEST_CURRENT_STATE, TEST_ACTION, TEST_NEXT_STATE)
    assert expected[0] == actual[0]


def test_running_inventory_penalty_per_step_reward():
    pnl = O.reward_eval(cfg_for(_abi.MBT_REW_PNL), TEST_CURRENT_STATE, TEST_ACTION, TEST_NEXT_STATE)
    expected = pnl - PHI * STEP_SIZE * abs(TEST_NEXT_STATE[:, 1]) ** 2
    actual = O.reward_eval(cfg_for(_abi.MBT_REW_RUNNING_INVENTORY_PENALTY, PHI, ALPHA), TEST_CURRENT_STATE,
                           TEST_ACTION, TEST_NEXT_STATE)
    assert actual[0] == pytest.approx(expected[0], abs=1e-5)


def _telescoping(observations, actions, start=0):
    mm, rip = cfg_for(_abi.MBT_REW_CJ_MM, PHI, ALPHA), cfg_for(_abi.MBT_REW_RUNNING_INVENTORY_PENALTY, PHI, ALPHA)
    q0 = observations[start][0, 1]
    L = TERMINAL_TIME - observations[start][0, 2]  # CjMmCriterion.reset  RewardFunctions.py:111-113
    tot_mm = tot_rip = 0.0
    for i in range(len(actions[start:])):
        cur, nxt, act = observations[start + i], observations[start + i + 1], actions[start + i]
        term = bool(nxt[0, 2] == 1)
        tot_mm += O.reward_eval(mm, cur, act, nxt, term, q0=q0, episode_length=L)[0]
        tot_rip += O.reward_eval(rip, cur, act, nxt, term)[0]
    return tot_mm, tot_rip


def test_cjmm_agrees_with_non_deconstructed_version():
    a, b = _telescoping(MOCK_OBSERVATIONS, MOCK_ACTIONS)
    assert a == pytest.approx(b, abs=1e-5)


def test_cjmm_agrees_with_nonzero_initial_inventory():
    obs = deepcopy(MOCK_OBSERVATIONS)
    obs[0][:, 1], obs[0][:, 0] = 2, -100
    obs[-1] = deepcopy(obs[-2])
    obs[-1][:, 2] = 1.0
    a, b = _telescoping(obs, MOCK_ACTIONS)
    assert a == pytest.approx(b, abs=1e-5)


def test_cjmm_agrees_on_partial_trajectory():
    a, b = _telescoping(MOCK_OBSERVATIONS, MOCK_ACTIONS, start=2)
    assert a == pytest.approx(b, abs=1e-5)
