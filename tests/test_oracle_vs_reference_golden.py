"""The oracle (oracle/mbt_oracle.c) against the committed outputs of the UNMODIFIED reference (tests/golden, made by
tools/make_golden.py with the reference's RNGs fed from the Philox draw contract): float64, bit-for-bit."""
import numpy as np
import pytest

from mbt_gym_b200 import _abi
from oracle import oracle as O
from tests.helpers import Golden, assert_same, golden_names


@pytest.mark.parametrize("name", golden_names())
def test_oracle_f64_equals_reference_fixture(name):
    g = Golden(name)
    orc = O.OracleEnv(g.config(_abi.MBT_F64))
    orc.seed(g.seed)
    spe = g.steps_per_episode
    for ep in range(g.n_episodes):
        assert_same(orc.reset(), g.reset_obs[ep], exact=True, what=f"{name} reset obs")
        for k in range(ep * spe, (ep + 1) * spe):
            o, r, d = orc.step(g.actions[k])
            assert_same(o, g.obs[k], exact=g.exact, what=f"{name} obs step {k}")
            assert_same(r, g.rew[k], exact=g.exact, what=f"{name} rew step {k}")
            assert d == g.done[k]
    assert_same(orc.state, g.final_state, exact=g.exact, what=f"{name} final state")


@pytest.mark.parametrize("name", ["as_pnl", "hawkes_pnl", "oe_ou_cjoe"])
def test_oracle_f32_tracks_f64(name):
    """float32 mode = the same formulas in float: early steps agree to float accuracy, discrete events identical."""
    g = Golden(name)
    a, b = O.OracleEnv(g.config(_abi.MBT_F64)), O.OracleEnv(g.config(_abi.MBT_F32))
    for e in (a, b):
        e.seed(g.seed)
        e.reset()
    for k in range(5):
        o64, r64, _ = a.step(g.actions[k])
        o32, r32, _ = b.step(g.actions[k])
        assert np.array_equal(o64[:, 1], o32[:, 1].astype(np.float64)) or name == "oe_ou_cjoe"  # inventories
        np.testing.assert_allclose(o32, o64, rtol=2e-6, atol=2e-3)


@pytest.mark.parametrize("name,cols", [("as_pnl_normalised", [1, 2]), ("hawkes_pnl", [1, 3, 5]), ("oe_ou_cjoe", [0, 4])])
def test_oracle_column_select_equals_reference_columns(name, cols):
    """ReduceStateSizeWrapper fused into the observation store = the reference's observation with those columns kept."""
    g = Golden(name)
    cfg = g.config(_abi.MBT_F64, obs_select=sum(1 << c for c in cols))
    orc = O.OracleEnv(cfg)
    orc.seed(g.seed)
    assert_same(orc.reset(), g.reset_obs[0][:, cols], what=f"{name} reset")
    for k in range(10):
        o, r, _ = orc.step(g.actions[k])
        assert_same(o, g.obs[k][:, cols], what=f"{name} obs {k}")
        assert_same(r, g.rew[k], what=f"{name} rew {k}")
