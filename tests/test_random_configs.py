"""Randomised configurations (seeded): model kinds, parameters, sizes, normalisation flags, start times and initial
inventories drawn at random within what the path supports.

  * `reference` (build container): the facade flattens to the same mbt_config bytes the reference's objects give, and the
    oracle equals the unmodified reference step-for-step under injected draws -- bit-for-bit;
  * `gpu` (B200 box): the CUDA path through the public API equals the oracle bit-for-bit in float64 and float32.
"""
import numpy as np
import pytest

from mbt_gym_b200 import _abi
from oracle import oracle as O
from tests.helpers import assert_same, build_facade_env


def random_spec(rng):
    dyn = rng.choice(["limit", "limit", "touch", "limit_and_market", "speed"])
    T = float(rng.choice([0.5, 1.0, 2.0]))
    spec = dict(N=int(rng.integers(1, 70)), n_steps=int(rng.integers(3, 13)), terminal_time=T,
                seed=int(rng.integers(1, 2 ** 31)), dynamics=dyn)
    mk = rng.choice(["bm", "bm", "gbm", "ou", "constant"] + ([] if dyn == "speed" else ["bm_jump", "ou_jump"]))
    S0 = float(rng.uniform(20, 200))
    if mk == "bm":
        spec["midprice"] = dict(kind="bm", drift=float(rng.uniform(-1, 1)), volatility=float(rng.uniform(0.2, 3)), initial_price=S0)
    elif mk == "gbm":
        spec["midprice"] = dict(kind="gbm", drift=float(rng.uniform(-0.2, 0.2)), volatility=float(rng.uniform(0.05, 0.4)), initial_price=S0)
    elif mk == "ou":
        spec["midprice"] = dict(kind="ou", level=S0 + float(rng.uniform(-2, 2)), speed=float(rng.uniform(0.01, 1.0)),
                                volatility=float(rng.uniform(0.2, 3)), initial_price=S0)
    elif mk == "bm_jump":
        spec["midprice"] = dict(kind="bm_jump", drift=float(rng.uniform(-1, 1)), volatility=float(rng.uniform(0.2, 3)),
                                jump=float(rng.uniform(0.05, 1.0)), initial_price=S0)
    elif mk == "ou_jump":
        spec["midprice"] = dict(kind="ou_jump", level=S0 + float(rng.uniform(-2, 2)), speed=float(rng.uniform(0.01, 1.0)),
                                volatility=float(rng.uniform(0.2, 3)), jump=float(rng.uniform(0.05, 1.0)), initial_price=S0)
    else:
        spec["midprice"] = dict(kind="constant", initial_price=S0)
    rewards = ["pnl", "rip", "cjmm", "exputil"]
    if dyn == "speed":
        ik = rng.choice(["temp_perm", "temp_perm", "temp_power", "temp_transient", "transient"])
        tr = dict(transient=float(rng.uniform(0.1, 1.0)), resilience=float(rng.uniform(0.1, 3.0)),
                  initial=float(rng.uniform(0, 0.05)), kernel=float(rng.uniform(0.1, 1.0)))
        spec["impact"] = {"temp_perm": dict(kind="temp_perm", temp=float(rng.uniform(0.001, 0.05)), perm=float(rng.uniform(0.001, 0.05))),
                          "temp_power": dict(kind="temp_power", temp=float(rng.uniform(0.001, 0.05)), exponent=1.0),
                          "temp_transient": dict(kind="temp_transient", temp=float(rng.uniform(0.001, 0.05)), **tr),
                          "transient": dict(kind="transient", **tr)}[str(ik)]
        rewards = ["pnl", "rip", "cjoe", "cjmm"]
        spec["initial_inventory"] = int(rng.integers(-50, 51))
    else:
        ak = rng.choice(["poisson", "poisson", "poisson_nonlinear", "hawkes"])
        if ak == "hawkes":
            spec["arrival"] = dict(kind="hawkes", baseline=[float(rng.uniform(5, 60)), float(rng.uniform(5, 60))],
                                   jump=float(rng.uniform(5, 60)), speed=float(rng.uniform(10, 90)))
        else:
            spec["arrival"] = dict(kind=str(ak), intensity=[float(rng.uniform(20, 400)), float(rng.uniform(20, 400))])
        if dyn != "touch":
            spec["fill"] = dict(kind="exp", fill_exponent=float(rng.uniform(0.5, 3.0)))
        else:
            spec["half_spread"] = float(rng.uniform(0.1, 1.0))
        if dyn == "limit_and_market":
            spec["half_spread"] = float(rng.uniform(0.1, 1.0))
        spec["initial_inventory"] = [-3, 4] if rng.random() < 0.3 else int(rng.integers(-3, 4))
    rk = str(rng.choice(rewards))
    spec["reward"] = {"pnl": dict(kind="pnl"), "rip": dict(kind="rip", phi=float(rng.uniform(0, 0.1)), alpha=float(rng.uniform(0, 1))),
                      "cjmm": dict(kind="cjmm", phi=float(rng.uniform(0, 0.1)), alpha=float(rng.uniform(0, 0.1))),
                      "cjoe": dict(kind="cjoe", phi=float(rng.uniform(0, 0.1)), alpha=float(rng.uniform(0, 0.1))),
                      "exputil": dict(kind="exputil", risk_aversion=float(rng.uniform(0.001, 0.02)))}[rk]
    spec["max_inventory"] = int(rng.choice([2, 5, 100, 10_000]))
    if rng.random() < 0.2:
        spec["max_cash"] = float(rng.uniform(100, 1000))
    if rng.random() < 0.3:
        spec["start_time"] = float(rng.uniform(0, 0.6 * T))
    spec["initial_cash"] = float(rng.choice([0.0, 100.0]))
    spec["normalise_obs"] = bool(rng.random() < 0.4)
    spec["normalise_action"] = bool(rng.random() < 0.4) and dyn != "touch"  # MultiBinary has no bounds to normalise
    return spec


def random_fill(spec, rng):
    """Swap the exponential fill function of some limit-order specs for a batch-reduced one (separate generator, so the
    sequence of specs drawn by `random_spec` stays what it was)."""
    pick = rng.choice(["exp", "exp", "triangular", "power", "exogenous"])
    if "fill" not in spec or pick == "exp":
        return spec
    if pick == "exogenous":
        spec["fill"] = dict(kind="exogenous", fill_exponent=float(rng.uniform(0.5, 3.0)), base=float(rng.uniform(0.3, 1.0)),
                            best_depths=[float(rng.uniform(0.0, 1.0)), float(rng.uniform(0.0, 1.0))])
        return spec
    if pick == "triangular":
        mfd = float(rng.uniform(0.5, 2.0))
        spec["fill"] = dict(kind="triangular", max_fill_depth=mfd)
        spec["depth_range"] = [-0.2, 1.3 * mfd]
    else:
        spec["fill"] = dict(kind="power", fill_exponent=float(rng.choice([1.0, 1.5, 2.0, 2.5])),
                            fill_multiplier=float(rng.uniform(0.5, 2.0)))
        spec["depth_range"] = [0.0, 3.0]
    return spec


def random_heston(spec, rng):
    """Swap some (geometric) Brownian midprices for the Heston model (separate generator, like `random_fill`)."""
    heston = rng.random() < 0.35
    pars = dict(kind="heston", drift=float(rng.uniform(-0.2, 0.2)), speed=float(rng.uniform(0.5, 5.0)),
                level=float(rng.uniform(0.01, 0.3)), corr=float(rng.choice([-0.8, -0.3, 0.0, 0.5, 1.0])),
                volvol=float(rng.uniform(0.1, 1.5)), initial_variance=float(rng.uniform(0.0, 0.4)))
    nine_columns = spec.get("fill", {}).get("kind") == "exogenous" and spec.get("arrival", {}).get("kind") == "hawkes"
    if heston and spec["midprice"]["kind"] in ("bm", "gbm") and not nine_columns:  # at most 8 observation columns
        spec["midprice"] = dict(pars, initial_price=spec["midprice"]["initial_price"])
        spec["normalise_obs"] = False  # undefined for Heston in the reference (bounds for one of two columns)
    return spec


def random_specs(n, seed):
    rng, rng_fill, rng_mid = np.random.default_rng(seed), np.random.default_rng(seed + 1), np.random.default_rng(seed + 2)
    return [random_heston(random_fill(random_spec(rng), rng_fill), rng_mid) for _ in range(n)]


@pytest.mark.reference
def test_random_configs_oracle_and_facade_vs_live_reference():
    from oracle import ref_shim as R

    exact = inexact = 0
    for spec in random_specs(40, 20260925):
        out = R.run_pair(spec, n_steps_run=spec["n_steps"], n_episodes=2)
        # numpy goes through libm for exp() / non-trivial pow(): decisions stay identical, values agree to ~1e-11
        libm = spec["reward"]["kind"] == "exputil"
        for key in ("obs", "rew", "reset"):
            assert_same(out["orc_" + key], out["ref_" + key], exact=not libm, what=f"{spec} {key}")
        assert out["ref_done"] == out["orc_done"]
        exact += not libm
        inexact += libm
        cfg = build_facade_env(spec)._build_config()
        assert bytes(cfg) == bytes(out["cfg"]), f"facade config differs from the reference's for {spec}"
    assert exact >= 25


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["float64", "float32"])
def test_random_configs_gpu_equals_oracle(precision):
    from oracle import ref_shim as R  # only for make_actions (no reference needed)

    prec = _abi.MBT_F64 if precision == "float64" else _abi.MBT_F32
    dt = np.float64 if precision == "float64" else np.float32
    for spec in random_specs(60, 424242):
        env = build_facade_env(spec, precision=precision, copy_outputs=True)
        cfg = env._build_config()
        orc = O.OracleEnv(cfg)
        orc.seed(spec["seed"])
        acts = R.make_actions(spec, env, spec["n_steps"] * 2, 11).astype(dt)
        k = 0
        for _ep in range(2):
            assert_same(env.reset(), orc.reset(), what=f"{spec} reset")
            for _ in range(spec["n_steps"]):
                o, r, d, _ = env.step(acts[k])
                oo, rr, dd = orc.step(acts[k])
                assert_same(o, oo, what=f"{spec} obs step {k}")
                assert_same(r, rr, what=f"{spec} rew step {k}")
                assert bool(d[0]) == dd
                k += 1
        assert_same(env.state, orc.state, what=f"{spec} final state")
        env.close()
