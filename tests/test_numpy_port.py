"""The NumPy port that `bench.py` times as the CPU baseline (oracle/numpy_port.py) against the LIVE, unmodified reference
under shared PCG64 seeds: every branch the bench can select (`--workload as|cjmm|hawkes|oe`) reproduces the reference's
observations, rewards and done flags over a whole episode.  Needs /root/reference (build container only; the port's AS
branch is additionally pinned to the notebook golden in tests/test_reference_live.py)."""
import numpy as np
import pytest

from oracle import numpy_port as P
from oracle import ref_shim as R

pytestmark = pytest.mark.skipif(not R.reference_available(), reason="needs /root/reference (build container only)")

AS = dict(midprice=dict(kind="bm", volatility=2.0, initial_price=100.0),
          arrival=dict(kind="poisson", intensity=[140.0, 140.0]), fill=dict(kind="exp", fill_exponent=1.5))
N, SEED = 257, 4321

# the same markets bench.make_env / numpy_port.make_env build (SURVEY.md 8d synthetic inputs)
SPECS = {
    "as": dict(N=N, n_steps=200, terminal_time=1.0, dynamics="limit", reward=dict(kind="pnl"), max_inventory=200, **AS),
    "cjmm": dict(N=N, n_steps=200, terminal_time=1.0, dynamics="limit", reward=dict(kind="cjmm", phi=0.01, alpha=0.001),
                 max_inventory=100, **AS),
    "hawkes": dict(N=N, n_steps=200, terminal_time=1.0, dynamics="limit", reward=dict(kind="pnl"), max_inventory=200,
                   midprice=AS["midprice"], fill=AS["fill"],
                   arrival=dict(kind="hawkes", baseline=[10.0, 10.0], jump=40.0, speed=60.0)),
    "oe": dict(N=N, n_steps=200, terminal_time=1.0, dynamics="speed",
               midprice=dict(kind="ou", level=100.0, speed=1.0, volatility=2.0, initial_price=100.0),
               impact=dict(kind="temp_perm", temp=0.01, perm=0.01), reward=dict(kind="cjoe", phi=0.01, alpha=0.001),
               initial_inventory=100, max_inventory=10_000),
}


@pytest.mark.parametrize("workload", ["as", "cjmm", "hawkes", "oe"])
@pytest.mark.parametrize("actions", ["fixed", "random"])
def test_port_equals_live_reference_under_a_shared_seed(workload, actions):
    ref = R.build_reference_env(SPECS[workload])
    ref.seed(SEED)  # process i gets SEED + i + 1 (TradingEnvironment.py:345-348), like the port's three generators
    port = P.make_env(workload, N, seed=SEED)
    assert np.array_equal(ref.reset(), port.reset())
    rng = np.random.default_rng(9)
    for k in range(200):
        if actions == "fixed":
            a = P.fixed_action(workload, N)
        elif workload == "oe":
            a = rng.uniform(-10, 10, size=(N, 1))
        else:
            a = rng.uniform(0.0, 3.0, size=(N, 2))
        o1, r1, d1, _ = ref.step(a.copy())
        o2, r2, d2, _ = port.step(a.copy())
        assert np.array_equal(np.asarray(o1, float), o2), f"{workload} obs step {k}"
        np.testing.assert_allclose(np.asarray(r1, float), r2, rtol=1e-13, atol=1e-11, err_msg=f"{workload} rewards step {k}")
        assert bool(d1[0]) == bool(d2[0]) == (k == 199)
